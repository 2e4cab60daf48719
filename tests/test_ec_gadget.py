"""SURVEY 8 f-4 pinned against the reference: the round model of oracle/ec_rounds.py (which the CUDA kernels are
tested against bit for bit in test_ec_rounds.py) must agree, error for error, with the instruction stream the
UNMODIFIED reference emits for CSSCode.error_correct (css_code.py:436-470), recorded in
tests/golden/ec_gadget_golden.json by oracle/gen_ec_gadget_golden.py.

The interpreter below knows nothing about the model: it is a Pauli-frame simulator for exactly the instructions
in that stream, vectorised over shots.
  CNOT c t          X_t ^= X_c, Z_c ^= Z_t
  H q               X_q <-> Z_q
  MEASURE q m       m = (ideal outcome of q) ^ X_q     -- the ideal outcomes of a block are a codeword of the
                    classical code the reference names (css_code.py:24-27: C_2 in the Z basis, C_1 in the X basis)
  MOVE/AND/XOR/IOR/NOT   the reference's classical decoder text, run by quil_text.run (tested against the
                    reference's own text in test_quil_text.py)
  while-block       css_code.encode_plus / encode_zero (css_code.py:314-366): verified preparation of ancilla
                    block ``a`` with helper ``b``; taken as ideal, followed by the injected ancilla error
"""

import json
import os

import numpy as np
import pytest

from oracle import css as ocss, ec_rounds as oec, gf2 as ogf2
from quantum_css_codes_b200 import codes, quil_text

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "ec_gadget_golden.json")
CLASSICAL = ("MOVE", "AND", "XOR", "IOR", "NOT")


class Frame:
    def __init__(self, n, scratch, shots):
        self.x = {f"{t}{i}": np.zeros(shots, dtype=np.uint8) for t in "dab" for i in range(n)}
        self.z = {q: np.zeros(shots, dtype=np.uint8) for q in self.x}
        self.mem = {f"{t}{k}": np.zeros((n, shots), dtype=np.uint8) for t in "dab" for k in "xz"}
        self.mem["scratch"] = np.zeros((scratch, shots), dtype=np.uint8)
        self.n = n

    def set_block(self, tag, ex, ez):
        for i in range(self.n):
            self.x[f"{tag}{i}"] = ex[:, i].copy()
            self.z[f"{tag}{i}"] = ez[:, i].copy()

    def block(self, tag):
        return (np.stack([self.x[f"{tag}{i}"] for i in range(self.n)], axis=1),
                np.stack([self.z[f"{tag}{i}"] for i in range(self.n)], axis=1))


def run_gadget(program, fr, ancilla_errors, ideal_words):
    """Execute one recorded error_correct gadget on the Pauli frame ``fr``.  ancilla_errors: the (x, z) error of
    block ``a`` after each of its two preparations; ideal_words: the noiseless outcomes of the two measured words."""
    prepared, measured, pending = 0, 0, []
    in_measure = False

    def flush():
        if pending:
            quil_text.run(pending, fr.mem)
            pending.clear()

    for ins in program:
        if isinstance(ins, dict):
            flush()
            assert "while" in ins, "only the preparation loops are blocks at the top level"
            zero = np.zeros_like(ancilla_errors[0][0])
            fr.set_block("b", zero, zero)
            fr.set_block("a", *ancilla_errors[prepared])
            for name in ("ax", "az", "bx", "bz"):
                fr.mem[name][...] = 0
            prepared += 1
            continue
        op, *args = ins.split()
        if op in CLASSICAL:
            if in_measure:
                in_measure = False
                measured += 1
            pending.append(ins)
            continue
        flush()
        if op == "CNOT":
            c, t = args
            fr.x[t] ^= fr.x[c]
            fr.z[c] ^= fr.z[t]
        elif op == "H":
            q = args[0]
            fr.x[q], fr.z[q] = fr.z[q], fr.x[q]
        elif op == "MEASURE":
            q, cell = args
            region, _, rest = cell.partition("[")
            fr.mem[region][int(rest[:-1])] = ideal_words[measured][:, int(q[1:])] ^ fr.x[q]
            in_measure = True
        else:
            raise AssertionError(f"unexpected instruction {ins}")
    flush()
    assert prepared == 2 and measured == 2


def random_codewords(rng, h, shots):
    """Uniform elements of the null space of h, (shots, n)."""
    basis = np.array(ogf2.null_space(np.array(h)), dtype=np.uint8)
    if basis.size == 0:
        return np.zeros((shots, np.array(h).shape[1]), dtype=np.uint8)
    coeff = rng.integers(0, 2, size=(shots, basis.shape[0]), dtype=np.uint8)
    return (coeff @ basis) & 1


def depolarising(rng, shots, n, p):
    kind = rng.choice(4, size=(shots, n), p=[1 - p, p / 3, p / 3, p / 3])      # I, X, Y, Z
    return ((kind == 1) | (kind == 2)).astype(np.uint8), ((kind == 3) | (kind == 2)).astype(np.uint8)


def _compare(name, p, round_update, shots=3000):
    """Runs two rounds through the recorded gadget and through ``round_update``; returns True when they agree."""
    with open(GOLDEN) as fh:
        gold = json.load(fh)[name]
    n, program = gold["n"], gold["program"]
    code = ocss.build_css(*[np.array(h) for h in getattr(codes, name)()])
    assert code.n == n
    rng = np.random.default_rng(7 + n)
    fr = Frame(n, gold["scratch"], shots)
    # model state: physical error e and frame f, both starting from something non-trivial
    e_x, e_z = depolarising(rng, shots, n, p)
    f_x, f_z = depolarising(rng, shots, n, p / 2)
    fr.set_block("d", e_x, e_z)
    fr.mem["dx"][...] = f_x.T
    fr.mem["dz"][...] = f_z.T
    for _ in range(2):                                   # two consecutive rounds: the frame carries over
        d = depolarising(rng, shots, n, p)
        a = depolarising(rng, shots, n, p)
        b = depolarising(rng, shots, n, p)
        # gadget side: this round's data error lands on the data block, then the recorded program runs
        gx, gz = fr.block("d")
        fr.set_block("d", gx ^ d[0], gz ^ d[1])
        words = [random_codewords(rng, code.parity_check_c2, shots),       # Z-basis measurement: a word of C_2
                 random_codewords(rng, code.parity_check_c1, shots)]       # X-basis measurement: a word of C_1
        run_gadget(program, fr, [a, b], words)
        # model side
        round_update(code, e_x, e_z, f_x, f_z, d, a, b)
        gx, gz = fr.block("d")
        if not (np.array_equal(gx, e_x) and np.array_equal(gz, e_z) and
                np.array_equal(fr.mem["dx"].T, f_x) and np.array_equal(fr.mem["dz"].T, f_z)):
            return False
    assert f_x.any() and f_z.any() and (e_x ^ f_x).any()
    return True


@pytest.mark.parametrize("name", ["steane", "shor9"])
@pytest.mark.parametrize("p", [0.02, 0.15])
def test_round_model_equals_the_reference_gadget(name, p):
    assert _compare(name, p, oec.round_update)


def _mutants():
    """Wrong models the comparison must reject: each changes ONE thing the gadget fixes."""
    from oracle import montecarlo as omc

    def make(drop_back_action=False, swap_sides=False, bx_early=False):
        def update(code, e_x, e_z, f_x, f_z, d, a, b):
            h2, t2, lz = ocss.pauli_side(code, 1 if swap_sides else 2)
            h1, t1, lx = ocss.pauli_side(code, 2 if swap_sides else 1)
            e_x ^= d[0]
            e_z ^= d[1]
            if not drop_back_action:
                e_z ^= a[1]
            if bx_early:
                e_x ^= b[0]
            f_x ^= omc.decode_batch(h2, t2, lz, e_x ^ a[0] ^ f_x)["corr"].astype(np.uint8)
            if not bx_early:
                e_x ^= b[0]
            f_z ^= omc.decode_batch(h1, t1, lx, e_z ^ b[1] ^ f_z)["corr"].astype(np.uint8)
        return update
    return {"no_z_back_action": make(drop_back_action=True), "sides_swapped": make(swap_sides=True),
            "ancilla_b_x_before_the_x_measurement": make(bx_early=True)}


@pytest.mark.parametrize("mutant", ["no_z_back_action", "sides_swapped", "ancilla_b_x_before_the_x_measurement"])
def test_the_comparison_rejects_wrong_models(mutant):
    """The pin has teeth: a model without the CNOT's Z back-action, with the two sides' matrices exchanged (only a code
    whose sides differ can tell: Shor-9), or with the |0>_L ancilla's X errors reaching the data before the X measurement
    does NOT reproduce the reference's gadget."""
    assert not _compare("shor9", 0.02, _mutants()[mutant])
    assert _compare("shor9", 0.02, oec.round_update)


def test_golden_is_the_gadget_of_css_code_error_correct():
    """Shape of the recorded stream: prepare |+>_L, CNOT data -> a, measure, decode with c2; prepare |0>_L,
    CNOT a -> data, H, measure, decode with c1 (css_code.py:456-470)."""
    with open(GOLDEN) as fh:
        gold = json.load(fh)
    for name, g in gold.items():
        n = g["n"]
        quantum = [i for i in g["program"] if isinstance(i, str) and i.split()[0] not in CLASSICAL]
        want = [f"CNOT d{i} a{i}" for i in range(n)] + [f"MEASURE a{i} scratch[{i}]" for i in range(n)] + \
               [f"CNOT a{i} d{i}" for i in range(n)] + [f"H a{i}" for i in range(n)] + \
               [f"MEASURE a{i} scratch[{i}]" for i in range(n)]
        assert quantum == want, name
        assert sum(isinstance(i, dict) for i in g["program"]) == 2
