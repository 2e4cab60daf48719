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


# ---- tile-major batches (sparse any-size syndrome path, qcss_syndrome_tiles) ---------------------------------
# [tile][plane][16 uint64]: the n plane rows of one tile of 1024 shots are contiguous, so a kernel stages a
# tile (or a contiguous run of its planes) with one bulk copy; the last tile is padded with zero bits.
TILE_SHOTS = 1024
TILE_WORDS = TILE_SHOTS // 64


def pack_tiles(vectors):
    """(shots, n) 0/1 array -> (ceil(shots / 1024), n, 16) uint64 tiles."""
    vectors = np.asarray(vectors)
    if vectors.ndim != 2:
        raise ValueError("expected a (shots, n) array")
    shots, n = vectors.shape
    tiles = (shots + TILE_SHOTS - 1) // TILE_SHOTS
    bits = np.zeros((n, tiles * TILE_SHOTS), dtype=np.uint8)
    bits[:, :shots] = (vectors.T & 1)
    words = np.packbits(bits, axis=1, bitorder="little").view(np.uint64).reshape(n, tiles, TILE_WORDS)
    return np.ascontiguousarray(words.transpose(1, 0, 2))


def unpack_tiles(tiles, shots):
    """(tiles, rows, 16) uint64 -> (shots, rows) uint8."""
    tiles = np.ascontiguousarray(tiles, dtype=np.uint64)
    count, rows, _ = tiles.shape
    words = np.ascontiguousarray(tiles.transpose(1, 0, 2)).reshape(rows, count * TILE_WORDS)
    bits = np.unpackbits(words.view(np.uint8), axis=1, bitorder="little")
    return np.ascontiguousarray(bits[:, :shots].T)


def planes_to_tiles(planes, shots):
    """(n, stride) plane-major words -> tile-major (same bits)."""
    planes = np.ascontiguousarray(planes, dtype=np.uint64)
    n = planes.shape[0]
    tiles = (shots + TILE_SHOTS - 1) // TILE_SHOTS
    padded = np.zeros((n, tiles * TILE_WORDS), dtype=np.uint64)
    take = min(planes.shape[1], tiles * TILE_WORDS)
    padded[:, :take] = planes[:, :take]
    return np.ascontiguousarray(padded.reshape(n, tiles, TILE_WORDS).transpose(1, 0, 2))


def events_from_arrays(x_errors, z_errors, first_shot=0):
    """Sparse form of a batch: uint64 events ``shot << 18 | qubit << 2 | pauli`` (pauli 1 = X, 2 = Z, 3 = Y), sorted by
    shot then qubit -- the input of ``qcss_decode_xz_sparse`` -- from two (shots, n) 0/1 arrays."""
    pauli = (np.asarray(x_errors) & 1).astype(np.uint64) | ((np.asarray(z_errors) & 1).astype(np.uint64) << np.uint64(1))
    shot, qubit = np.nonzero(pauli)
    return ((shot.astype(np.uint64) + np.uint64(first_shot)) << np.uint64(18)) | (qubit.astype(np.uint64) << np.uint64(2)) \
        | pauli[shot, qubit]


def arrays_from_events(events, shots, n, first_shot=0):
    """Inverse of ``events_from_arrays`` (repeated events compose by XOR): two (shots, n) uint8 arrays."""
    events = np.asarray(events, dtype=np.uint64)
    shot = (events >> np.uint64(18)).astype(np.int64) - first_shot
    qubit = ((events >> np.uint64(2)) & np.uint64(0xFFFF)).astype(np.int64)
    ex = np.zeros((shots, n), dtype=np.uint8)
    ez = np.zeros((shots, n), dtype=np.uint8)
    np.bitwise_xor.at(ex, (shot, qubit), (events & np.uint64(1)).astype(np.uint8))
    np.bitwise_xor.at(ez, (shot, qubit), ((events >> np.uint64(1)) & np.uint64(1)).astype(np.uint8))
    return ex, ez
