"""
ORACLE (test infrastructure only -- never imported by the product package).

Pauli-frame Monte Carlo of repeated Steane error correction, restated in numpy in ERROR space: explicit
(shots, n) error and frame arrays, the reference's table decode for every measurement.  The device code
(csrc/ec_rounds.cuh) works in syndrome space instead; agreement of the two on identical Philox streams is
the parity check for SURVEY 8 f-4.

**What pins it.**  css_code.py only EMITS this gadget (CSSCode.error_correct, css_code.py:436-470) and
test/test_fidelity.py runs it on a QVM with a T1/T2 noise model; the reference has no Pauli noise model, so WHERE
errors enter (one depolarising layer on the data per round, one on each verified ancilla) is this model's choice.
Everything else is the reference's and is pinned against it: tests/test_ec_gadget.py takes the instruction stream the
UNMODIFIED reference emits for error_correct (tests/golden/ec_gadget_golden.json, recorded by
oracle/gen_ec_gadget_golden.py under a pyquil stub), propagates explicit Pauli errors through its CNOT / H / MEASURE
gates, runs its classical decoder text, and requires ``round_update`` below to give the same physical error and the
same frame registers, error for error, over two consecutive rounds, for a self-dual code (Steane) and one whose two
sides differ (Shor-9).  That covers: which errors copy where through the two transversal CNOTs, the order of the two
halves, which parity check / table serves which Pauli type (css_code.py:456-470: X errors <- parity_check_c2 /
_c2_syndromes, Z errors <- parity_check_c1 / _c1_syndromes), and the frame update of quil_classical_correct
(css_code.py:649-685): syndrome of (measured word ^ frame), frame ^= table.get(key, 0).

Model (per round r, streams = Philox site 32 * stream + qubit, oracle/philox.py):
  data      e ^= depolarising(p_data)                         stream 3r
  ancilla A a = depolarising(p_anc)                           stream 3r + 1   (|+>_L, css_code.py:345-366)
            CNOT data -> A:  e_z ^= a_z ; measured = e_x ^ a_x (mod the codeword, which has zero syndrome)
            f_x ^= table2.get(key(H2 . (measured ^ f_x)), 0)
  ancilla B b = depolarising(p_anc)                           stream 3r + 2   (|0>_L, css_code.py:314-343)
            CNOT B -> data:  e_x ^= b_x ; measured = e_z ^ b_z
            f_z ^= table1.get(key(H1 . (measured ^ f_z)), 0)
after the last round the residual e ^ f is decoded once more, noiselessly, and tallied like
montecarlo.tally_xz.
"""

import numpy as np

from . import css as ocss
from . import montecarlo as omc
from . import philox as ophilox


def _bits(planes, shots):
    n = planes.shape[0]
    b = np.unpackbits(np.ascontiguousarray(planes).view(np.uint8).reshape(n, -1), axis=1, bitorder="little")
    return np.ascontiguousarray(b[:, :shots].T)


def draw(seed, first_shot, shots, n, p, stream):
    """(x, z) error bits (shots, n) of one depolarising layer: Philox sites 32 * stream .. + n - 1."""
    assert first_shot % 32 == 0 and n <= 32
    ex, ez = ophilox.sample_words(seed, first_shot // 32, (shots + 31) // 32, n, p, site0=32 * stream)
    return _bits(ex, shots), _bits(ez, shots)


def round_update(code, e_x, e_z, f_x, f_z, d, a, b):
    """One round of the model on explicit (shots, n) bit arrays, in place: physical data error (e_x, e_z), frame
    (f_x, f_z); d, a, b = (x, z) pairs of this round's data / ancilla-A / ancilla-B errors.  The same update is what
    tests/test_ec_gadget.py obtains by propagating (e, a, b) through the instruction stream the UNMODIFIED reference
    emits for CSSCode.error_correct (tests/golden/ec_gadget_golden.json)."""
    h2, table2, lz = ocss.pauli_side(code, 2)          # X errors
    h1, table1, lx = ocss.pauli_side(code, 1)          # Z errors
    e_x ^= d[0]
    e_z ^= d[1]
    e_z ^= a[1]
    f_x ^= omc.decode_batch(h2, table2, lz, e_x ^ a[0] ^ f_x)["corr"].astype(np.uint8)
    e_x ^= b[0]
    f_z ^= omc.decode_batch(h1, table1, lx, e_z ^ b[1] ^ f_z)["corr"].astype(np.uint8)


def ec_rounds(code, p_data, p_ancilla, rounds, shots, seed=0, first_shot=0):
    """Tally dict of ``rounds`` rounds of Steane EC on ``shots`` shots for an oracle css code object."""
    n = code.n
    e_x = np.zeros((shots, n), dtype=np.uint8)
    e_z = np.zeros((shots, n), dtype=np.uint8)
    f_x = np.zeros((shots, n), dtype=np.uint8)
    f_z = np.zeros((shots, n), dtype=np.uint8)
    for r in range(rounds):
        round_update(code, e_x, e_z, f_x, f_z,
                     draw(seed, first_shot, shots, n, p_data, 3 * r),
                     draw(seed, first_shot, shots, n, p_ancilla, 3 * r + 1),
                     draw(seed, first_shot, shots, n, p_ancilla, 3 * r + 2))
    return omc.tally_xz(code, e_x ^ f_x, e_z ^ f_z)
