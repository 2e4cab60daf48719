"""
The reference's classical decoder as it goes over the wire: the straight-line classical Quil that
``css_code.quil_classical_correct`` / ``quil_classical_detect`` append to a program
(css_code.py:649-713, built from ``quil_classical.matmul`` / ``string_match`` / ``conditional_xor``,
quil_classical.py:60-111), emitted here as Quil TEXT without pyquil, plus a small interpreter for
exactly that instruction subset (MOVE / AND / XOR / IOR / NOT over bit registers), vectorised over
shots.  SURVEY 8 f-3: it ties the GPU lookup decoder to the on-wire form of the same decoder --
``tests`` run the emitted program on random frames and compare with the kernels' corrections.

Host-side only; nothing here is on the Monte-Carlo hot path.
"""

import numpy as np

from . import bin_matrix


class Chunk:
    """A slice ``name[start:end]`` of a classical memory region (quil_classical.py:10-53)."""

    def __init__(self, name, start, end):
        self.name, self.start, self.end = name, int(start), int(end)

    def __len__(self):
        return self.end - self.start

    def __getitem__(self, index):
        if isinstance(index, slice):
            start = 0 if index.start is None else index.start
            stop = len(self) if index.stop is None else index.stop
            if start < 0 or self.start + stop > self.end:
                raise IndexError("out of bounds")
            return Chunk(self.name, self.start + start, self.start + stop)
        if index < 0 or index >= len(self):
            raise IndexError("out of bounds")
        return f"{self.name}[{self.start + index}]"


def matmul(lines, mat, vec, result, scratch):
    """quil_classical.matmul (quil_classical.py:60-79): result = mat . vec over GF(2)."""
    m, n = mat.shape
    if len(vec) != n:
        raise ValueError("mat and vec are of incompatible sizes")
    if len(result) != m:
        raise ValueError("mat and result are of incompatible sizes")
    if len(scratch) < 1:
        raise ValueError("scratch buffer is too small")
    for i in range(m):
        lines.append(f"MOVE {result[i]} 0")
        for j in range(n):
            lines.append(f"MOVE {scratch[0]} {vec[j]}")
            lines.append(f"AND {scratch[0]} {int(mat[i][j])}")
            lines.append(f"XOR {result[i]} {scratch[0]}")


def string_match(lines, mem, vec, output, scratch):
    """quil_classical.string_match (quil_classical.py:81-97): output[0] = (mem == vec)."""
    n = len(mem)
    if np.asarray(vec).size != n:
        raise ValueError("length of mem and vec do not match")
    if len(scratch) < 1:
        raise ValueError("scratch buffer is too small")
    lines.append(f"MOVE {output[0]} 0")
    for i in range(n):
        lines.append(f"MOVE {scratch[0]} {mem[i]}")
        lines.append(f"XOR {scratch[0]} {int(vec[i])}")
        lines.append(f"IOR {output[0]} {scratch[0]}")
    lines.append(f"NOT {output[0]}")


def conditional_xor(lines, mem, vec, flag, scratch):
    """quil_classical.conditional_xor (quil_classical.py:99-111): mem ^= vec if flag[0]."""
    n = len(mem)
    if np.asarray(vec).size != n:
        raise ValueError("length of mem and vec do not match")
    for i in range(n):
        lines.append(f"MOVE {scratch[0]} {flag[0]}")
        lines.append(f"AND {scratch[0]} {int(vec[i])}")
        lines.append(f"XOR {mem[i]} {scratch[0]}")


def quil_classical_correct(codeword, errors, scratch, parity_check, syndromes):
    """css_code.quil_classical_correct (css_code.py:649-685) as a list of Quil lines."""
    parity_check = np.asarray(parity_check)
    m, n = parity_check.shape
    if len(codeword) != n:
        raise ValueError("codeword is of incorrect size")
    if len(errors) != n:
        raise ValueError("errors is of incorrect size")
    if len(scratch) < m + 2:
        raise ValueError("scratch buffer is too small")
    lines = []
    lines += [f"XOR {codeword[i]} {errors[i]}" for i in range(n)]
    syndrome = scratch[2:m + 2]
    matmul(lines, parity_check, codeword, syndrome, scratch[:2])
    lines += [f"XOR {codeword[i]} {errors[i]}" for i in range(n)]
    for key, correction in syndromes.items():
        match = bin_matrix.int_to_vec(key, m)
        matches = scratch[1:2]
        string_match(lines, syndrome, match, matches, scratch[:1])
        conditional_xor(lines, errors, correction, matches, scratch[:1])
    lines += [f"XOR {codeword[i]} {errors[i]}" for i in range(n)]
    return lines


def quil_classical_detect(codeword, errors, outcome, scratch, parity_check):
    """css_code.quil_classical_detect (css_code.py:687-713) as a list of Quil lines."""
    parity_check = np.asarray(parity_check)
    m, n = parity_check.shape
    if len(codeword) != n:
        raise ValueError("codeword is of incorrect size")
    if len(errors) != n:
        raise ValueError("errors is of incorrect size")
    if len(scratch) < m + 2:
        raise ValueError("scratch buffer is too small")
    lines = []
    lines += [f"XOR {codeword[i]} {errors[i]}" for i in range(n)]
    syndrome = scratch[2:m + 2]
    matmul(lines, parity_check, codeword, syndrome, scratch[:2])
    lines += [f"XOR {codeword[i]} {errors[i]}" for i in range(n)]
    lines.append(f"MOVE {outcome} 0")
    lines += [f"IOR {outcome} {syndrome[i]}" for i in range(m)]
    return lines


def run(lines, memory):
    """Execute classical Quil lines (MOVE / AND / XOR / IOR / NOT on BIT cells) on ``memory``:
    dict region name -> uint8 array of shape (cells, shots), updated in place.  Operands are
    ``name[index]`` cells or integer literals."""
    def cell(token):
        name, _, rest = token.partition("[")
        return memory[name][int(rest[:-1])]

    def value(token):
        return cell(token) if "[" in token else np.uint8(int(token) & 1)

    for line in lines:
        parts = line.split()
        op = parts[0]
        if op not in ("MOVE", "AND", "XOR", "IOR", "NOT"):
            raise ValueError(f"unsupported instruction: {line}")
        dst = cell(parts[1])
        if op == "MOVE":
            dst[...] = value(parts[2])
        elif op == "AND":
            dst &= value(parts[2])
        elif op == "XOR":
            dst ^= value(parts[2])
        elif op == "IOR":
            dst |= value(parts[2])
        else:
            dst ^= np.uint8(1)
    return memory
