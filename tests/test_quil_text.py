"""the classical decoder in its on-wire form (Quil text emitted by
css_code.quil_classical_correct / _detect, css_code.py:649-713) tied to the lookup decoder.

CPU part: the emitted text equals what the UNMODIFIED reference emits (tests/golden/
quil_classical_golden.json, made by oracle/gen_quil_golden.py), and the interpreter run over every
frame equals the oracle's table decode.  The reference's own quil_classical tests
(test/test_quil_classical.py:15-70) are restated against the interpreter instead of a QVM.
GPU part: the interpreted program equals the CUDA decoder (K1+K2) on the same frames.
"""

import json
import os

import numpy as np
import pytest

import css_code
from oracle import css as ocss, montecarlo as omc
from quantum_css_codes_b200 import codes, quil_text
from quantum_css_codes_b200.quil_text import Chunk

HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.fixture(scope="module")
def quil_golden():
    with open(os.path.join(HERE, "golden", "quil_classical_golden.json")) as fh:
        return json.load(fh)


def registers(m, n):
    return Chunk("codeword", 0, n), Chunk("errors", 0, n), Chunk("scratch", 0, m + 2)


def run_correct(h, table, frames, codewords=None):
    """Interpret the emitted corrector over a (shots, n) batch of error frames; returns the
    correction it folded into the frame and the codeword register afterwards."""
    h = np.asarray(h)
    m, n = h.shape
    shots = frames.shape[0]
    cw, er, sc = registers(m, n)
    prog = css_code.quil_classical_correct([], cw, er, sc, h, table)
    mem = {"codeword": np.zeros((n, shots), dtype=np.uint8) if codewords is None
           else np.ascontiguousarray(codewords.T.astype(np.uint8)),
           "errors": np.ascontiguousarray(frames.T.astype(np.uint8)),
           "scratch": np.ones((m + 2, shots), dtype=np.uint8)}          # dirty scratch on purpose
    quil_text.run(prog, mem)
    return (mem["errors"].T ^ frames.astype(np.uint8)), mem["codeword"].T


def run_detect(h, frames):
    h = np.asarray(h)
    m, n = h.shape
    shots = frames.shape[0]
    cw, er, sc = registers(m, n)
    prog = css_code.quil_classical_detect([], cw, er, "outcome[0]", sc, h)
    mem = {"codeword": np.zeros((n, shots), dtype=np.uint8),
           "errors": np.ascontiguousarray(frames.T.astype(np.uint8)),
           "scratch": np.ones((m + 2, shots), dtype=np.uint8),
           "outcome": np.ones((1, shots), dtype=np.uint8)}
    quil_text.run(prog, mem)
    assert np.array_equal(mem["errors"].T, frames)                       # detect leaves the frame alone
    return mem["outcome"][0]


# ---- emitted text against the unmodified reference -----------------------------------------------

@pytest.mark.parametrize("name,tag", [("steane", "c1"), ("steane", "c2"), ("qrm15", "c1")])
def test_emitted_text_equals_reference(quil_golden, name, tag):
    h1, h2 = getattr(codes, name)()
    code = ocss.build_css(np.array(h1), np.array(h2))
    h, table = ((code.parity_check_c1, code.c1_syndromes) if tag == "c1"
                else (code.parity_check_c2, code.c2_syndromes))
    m, n = h.shape
    cw, er, sc = registers(m, n)
    assert css_code.quil_classical_correct([], cw, er, sc, h, table) == quil_golden[f"{name}_{tag}_correct"]
    assert (css_code.quil_classical_detect([], cw, er, "outcome[0]", sc, h)
            == quil_golden[f"{name}_{tag}_detect"])


def test_emitter_argument_errors():
    """Size checks of css_code.py:660-665 / 698-703."""
    h = np.array(codes.steane()[0])
    cw, er, sc = registers(3, 7)
    with pytest.raises(ValueError, match="codeword is of incorrect size"):
        css_code.quil_classical_correct([], cw[0:6], er, sc, h, {})
    with pytest.raises(ValueError, match="errors is of incorrect size"):
        css_code.quil_classical_correct([], cw, er[0:6], sc, h, {})
    with pytest.raises(ValueError, match="scratch buffer is too small"):
        css_code.quil_classical_correct([], cw, er, sc[0:4], h, {})
    with pytest.raises(ValueError, match="scratch buffer is too small"):
        css_code.quil_classical_detect([], cw, er, "outcome[0]", sc[0:4], h)


def test_prog_with_iadd_receives_lines():
    """``prog`` may be anything with ``+=`` (a pyquil Program parses strings)."""
    class Recorder:
        def __init__(self):
            self.lines = []

        def __iadd__(self, line):
            assert isinstance(line, str)
            self.lines.append(line)
            return self

    h = np.array(codes.steane()[0])
    cw, er, sc = registers(3, 7)
    rec = css_code.quil_classical_detect(Recorder(), cw, er, "outcome[0]", sc, h)
    assert rec.lines == css_code.quil_classical_detect([], cw, er, "outcome[0]", sc, h)


# ---- the reference's quil_classical tests, on the interpreter ----------------------------------------

def test_matmul_like_reference():
    """test/test_quil_classical.py:15-40."""
    rng = np.random.default_rng(7)
    m, n = 20, 10
    mat = rng.integers(0, 2, size=(m, n))
    vec = rng.integers(0, 2, size=n)
    mem = Chunk("ro", 0, n + m + 1)
    lines = [f"MOVE {mem[i]} {int(vec[i])}" for i in range(n)]
    quil_text.matmul(lines, mat, mem[0:n], mem[n:n + m], mem[n + m:n + m + 1])
    ro = quil_text.run(lines, {"ro": np.zeros((n + m + 1, 1), dtype=np.uint8)})["ro"][:, 0]
    assert np.array_equal(ro[n:n + m], np.mod(np.matmul(mat, vec), 2))
    with pytest.raises(ValueError, match="mat and vec are of incompatible sizes"):
        quil_text.matmul([], mat, mem[0:n - 1], mem[n:n + m], mem[n + m:n + m + 1])
    with pytest.raises(ValueError, match="mat and result are of incompatible sizes"):
        quil_text.matmul([], mat, mem[0:n], mem[n:n + m - 1], mem[n + m:n + m + 1])


@pytest.mark.parametrize("vec1,vec2,expected", [
    ([0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 0], True),
    ([0, 0, 0, 0, 0, 0, 0, 1], [0, 0, 0, 0, 0, 0, 0, 1], True),
    ([0, 0, 0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 0, 0, 1, 1], True),
    ([0, 0, 0, 0, 0, 0, 0, 0], [0, 0, 0, 0, 0, 0, 0, 1], False),
    ([0, 0, 0, 0, 0, 0, 1, 0], [0, 0, 0, 0, 0, 0, 0, 1], False),
    ([0, 0, 0, 0, 0, 0, 1, 1], [0, 0, 0, 0, 0, 0, 0, 1], False),
])
def test_string_match_like_reference(vec1, vec2, expected):
    """test/test_quil_classical.py:42-70."""
    n = len(vec1)
    mem = Chunk("ro", 0, n + 2)
    lines = [f"MOVE {mem[i]} {vec2[i]}" for i in range(n)]
    quil_text.string_match(lines, mem[0:n], np.array(vec1), mem[n:n + 1], mem[n + 1:n + 2])
    ro = quil_text.run(lines, {"ro": np.zeros((n + 2, 1), dtype=np.uint8)})["ro"][:, 0]
    assert (ro[n] == 1) == expected


def test_chunk_like_reference():
    """test/test_quil_classical.py:114-150 (MemoryChunk)."""
    chunk = Chunk("test", 10, 20)
    assert (chunk.start, chunk.end, len(Chunk("test", 1, 10))) == (10, 20, 9)
    assert chunk[5] == "test[15]"
    sub = chunk[2:5]
    assert (sub.start, sub.end) == (12, 15)
    with pytest.raises(IndexError):
        chunk[10]
    with pytest.raises(IndexError):
        chunk[0:11]


def test_interpreter_rejects_other_instructions():
    with pytest.raises(ValueError, match="unsupported instruction"):
        quil_text.run(["H 0"], {})


# ---- the emitted decoder equals the table decode (oracle) -----------------------------------------

@pytest.mark.parametrize("name", ["steane", "qrm15"])
@pytest.mark.parametrize("which", [1, 2])
def test_emitted_decoder_equals_oracle(name, which):
    h1, h2 = getattr(codes, name)()
    code = ocss.build_css(np.array(h1), np.array(h2))
    h, table, lop = ocss.pauli_side(code, which)
    if name == "steane":
        frames = omc.all_patterns(code.n)                                # all 2^7 frames
    else:
        rng = np.random.default_rng(15 + which)
        frames = (rng.random((512, code.n)) < 0.12).astype(np.uint8)     # includes QRM table misses
    want = omc.decode_batch(h, table, lop, frames)
    corr, cw = run_correct(h, table, frames)
    assert np.array_equal(corr, want["corr"])
    assert np.array_equal(cw, frames ^ want["corr"])                     # zero codeword ^ corrected frame
    assert np.array_equal(run_detect(h, frames), want["synd"].any(axis=1).astype(np.uint8))


def test_emitted_decoder_on_nonzero_codeword():
    """The corrector acts on ``codeword ^ errors`` (css_code.py:667-676): with a real codeword in the
    register the correction is unchanged and the codeword comes back corrected."""
    h1, h2 = codes.steane()
    code = ocss.build_css(np.array(h1), np.array(h2))
    h, table, lop = ocss.pauli_side(code, 2)
    frames = omc.all_patterns(code.n)
    word = np.mod(np.array([1, 0, 1]) @ code.parity_check_c1, 2).astype(np.uint8)   # C_2-dual word: H2.word = 0
    assert not np.mod(h @ word, 2).any()
    words = np.broadcast_to(word, frames.shape)
    want = omc.decode_batch(h, table, lop, frames)
    corr, cw = run_correct(h, table, frames, codewords=words)
    assert np.array_equal(corr, want["corr"])
    assert np.array_equal(cw, words ^ frames ^ want["corr"])


# ---- the emitted decoder equals the CUDA decoder ------------------------------------------------------

@pytest.mark.gpu
@pytest.mark.parametrize("name,shots", [("steane", 4096), ("qrm15", 1024), ("golay23", 256)])
@pytest.mark.parametrize("which", [1, 2])
def test_emitted_decoder_equals_cuda_decoder(name, which, shots):
    h1, h2 = getattr(codes, name)()
    code = css_code.CSSCode(np.array(h1), np.array(h2))
    h, table = code._side(which)
    rng = np.random.default_rng(100 * which + code.n)
    frames = (rng.random((shots, code.n)) < 0.1).astype(np.uint8)
    frames[0] = 0
    out = code.decode(frames, which)
    corr, _ = run_correct(h, table, frames)
    assert np.array_equal(out["correction"], corr)
    detected = run_detect(h, frames)
    assert np.array_equal(code.syndromes(frames, which).any(axis=1).astype(np.uint8), detected)
