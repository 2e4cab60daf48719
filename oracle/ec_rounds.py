"""
ORACLE (test infrastructure only -- never imported by the product package).

Pauli-frame Monte Carlo of repeated Steane error correction, restated in numpy in ERROR space: explicit
(shots, n) error and frame arrays, the reference's table decode for every measurement.  The device code
(csrc/ec_rounds.cuh) works in syndrome space instead; agreement of the two on identical Philox streams is
the parity check for SURVEY 8 f-4.

**Parity unpinned by the reference**: css_code.py only EMITS this gadget (CSSCode.error_correct,
css_code.py:436-470) and test/test_fidelity.py runs it on a QVM with a T1/T2 noise model; there is no
Pauli model to compare with.  What IS the reference's: the frame update of every round, which follows
quil_classical_correct (css_code.py:649-685): syndrome of (measured word ^ frame) with the round's parity
check, frame ^= table.get(key, 0); and which matrices / tables serve which Pauli type (css_code.py:456-470:
X errors <- parity_check_c2 / _c2_syndromes, Z errors <- parity_check_c1 / _c1_syndromes).

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


def ec_rounds(code, p_data, p_ancilla, rounds, shots, seed=0, first_shot=0):
    """Tally dict of ``rounds`` rounds of Steane EC on ``shots`` shots for an oracle css code object."""
    n = code.n
    h2, table2, lz = ocss.pauli_side(code, 2)          # X errors
    h1, table1, lx = ocss.pauli_side(code, 1)          # Z errors
    e_x = np.zeros((shots, n), dtype=np.uint8)
    e_z = np.zeros((shots, n), dtype=np.uint8)
    f_x = np.zeros((shots, n), dtype=np.uint8)
    f_z = np.zeros((shots, n), dtype=np.uint8)
    for r in range(rounds):
        d_x, d_z = draw(seed, first_shot, shots, n, p_data, 3 * r)
        e_x ^= d_x
        e_z ^= d_z
        a_x, a_z = draw(seed, first_shot, shots, n, p_ancilla, 3 * r + 1)
        e_z ^= a_z
        f_x ^= omc.decode_batch(h2, table2, lz, e_x ^ a_x ^ f_x)["corr"].astype(np.uint8)
        b_x, b_z = draw(seed, first_shot, shots, n, p_ancilla, 3 * r + 2)
        e_x ^= b_x
        f_z ^= omc.decode_batch(h1, table1, lx, e_z ^ b_z ^ f_z)["corr"].astype(np.uint8)
    return omc.tally_xz(code, e_x ^ f_x, e_z ^ f_z)
