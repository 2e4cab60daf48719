"""
Bit-plane ("plane-major") batches: the device-native layout of error / syndrome batches.

A batch of ``shots`` binary vectors of length ``n`` is stored as ``n`` planes; plane ``j`` is a
little-endian bit array over shots: shot ``s`` is bit ``s % 64`` of ``uint64`` word ``s // 64``.
Planes are padded to a whole number of 512-bit groups so device kernels can use 128-bit loads on
any plane; padding bits are zero.  This is the transpose of the reference's shot-major
``(shots, n)`` int arrays (css_code.py:728 works on one length-n vector at a time).
"""

import numpy as np

WORD_ALIGN = 8          # plane stride is a multiple of 8 uint64 words (64 bytes)


def stride_words(shots):
    words = (int(shots) + 63) // 64
    return max(WORD_ALIGN, (words + WORD_ALIGN - 1) // WORD_ALIGN * WORD_ALIGN)


def pack_planes(vectors):
    """(shots, n) 0/1 array -> (n, stride_words(shots)) uint64 planes."""
    vectors = np.asarray(vectors)
    if vectors.ndim != 2:
        raise ValueError("expected a (shots, n) array")
    shots, n = vectors.shape
    stride = stride_words(shots)
    bits = np.zeros((n, stride * 64), dtype=np.uint8)
    bits[:, :shots] = (vectors.T & 1)
    return np.packbits(bits, axis=1, bitorder="little").view(np.uint64).reshape(n, stride)


def unpack_planes(planes, shots):
    """(n, stride) uint64 planes -> (shots, n) uint8."""
    planes = np.ascontiguousarray(planes, dtype=np.uint64)
    bits = np.unpackbits(planes.view(np.uint8), axis=1, bitorder="little")
    return np.ascontiguousarray(bits[:, :shots].T)


def unpack_plane(plane, shots):
    """(stride,) uint64 -> (shots,) uint8."""
    plane = np.ascontiguousarray(plane, dtype=np.uint64)
    return np.unpackbits(plane.view(np.uint8), bitorder="little")[:shots]
