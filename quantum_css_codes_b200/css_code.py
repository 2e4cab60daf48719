"""
Drop-in for the numeric core of the reference's ``css_code.py`` with the hot path on the GPU.

What is mirrored (same names, argument meaning, exceptions, evaluation order):
  CSSCode(parity_check_c1, parity_check_c2)           css_code.py:32-75
  CSSCode.n / k / t, r_1, r_2, parity_check_c1/2,
  _c1_syndromes / _c2_syndromes, is_transversal       css_code.py:65-96, 174-201
  z_operator_matrix / x_operator_matrix               css_code.py:124-136, 149-161
  syndrome_table, swap_columns, normalize_parity_check,
  codes_equal, is_doubly_even                         css_code.py:715-735, 783-785, 809-850
What is new (the Monte-Carlo hot path the north star adds):
  CSSCode.syndromes / decode / decode_xz / monte_carlo / sample_errors  -> CUDA kernels K1-K3
Host-side tie to the on-wire decoder:
  quil_classical_correct / quil_classical_detect      css_code.py:649-713 -- same signature; the
  instructions are appended to ``prog`` as Quil text lines (``quil_text.py``; a list, or anything with
  ``+=`` such as a pyquil Program, which parses strings), and ``quil_text.run`` interprets them.
Out of scope here (stays with the reference's host code): every method that
emits quantum gates into a pyquil Program (encode_*, error_correct, measure, ftqc rewriting).

Naming follows the reference's (inverted) convention, css_code.py:28-30 and 461-470:
  X errors <-> parity_check_c2 / _c2_syndromes / z_operator_matrix      ("which" = 2)
  Z errors <-> parity_check_c1 / _c1_syndromes / x_operator_matrix      ("which" = 1)
"""

import itertools

import numpy as np

from . import _native
from . import bin_matrix
from . import planes as _planes
from .errors import InvalidCodeError, UnsupportedGateError  # noqa: F401  (re-exported)

_LAYER_CHUNK = 1 << 16


def swap_columns(mat, indices):
    """Exchange two columns of ``mat`` in place (css_code.py:783-785)."""
    i, j = indices
    mat[:, [i, j]] = mat[:, [j, i]]


def normalize_parity_check(h, offset):
    """Bring ``h`` to standard form with an identity block at columns offset..offset+r-1.

    Mirrors css_code.py:809-836 step for step, including its side effects: ``h`` is updated in
    place with *un-reduced* integer row sums, qubit (column) swaps are chosen by the first odd
    entry of row i at or right of the diagonal when no row below can supply a pivot, and the
    result is ``(np.mod(h, 2), swaps)``.
    """
    r, n = h.shape
    if n < offset + r:
        raise ValueError("not enough columns")
    qubit_swaps = []
    for i in range(r):
        col = i + offset
        odd_below = np.flatnonzero(h[i:, col] % 2 == 1)
        if odd_below.size:
            if h[i, col] % 2 == 0:
                h[i, :] += h[i + int(odd_below[0]), :]
        else:
            odd_right = np.flatnonzero(h[i, col:] % 2 == 1)
            if odd_right.size == 0:
                raise InvalidCodeError("rows are not independent")
            qubit_swaps.append((col, col + int(odd_right[0])))
            swap_columns(h, qubit_swaps[-1])
        clear = h[:, col] % 2 == 1
        clear[i] = False
        h[clear, :] += h[i, :]
    return np.mod(h, 2), qubit_swaps


def normalize_parity_check_gpu(h, offset):
    """``normalize_parity_check`` on the device (``qcss_gf2_normalize``): the same
    ``(np.mod(h, 2), qubit_swaps)`` -- same pivot rule, same swap pairs in the same order, same
    exceptions -- computed by one CTA over bit-packed rows.  ``h`` is updated in place like the
    reference's, except that it receives the REDUCED matrix (entries 0/1) where the reference leaves
    un-reduced integer row sums of the same parity."""
    r, n = h.shape
    if n < offset + r:
        raise ValueError("not enough columns")
    bits = np.mod(np.asarray(h), 2).astype(np.uint8)
    out, swaps, status = _native.gf2_normalize_packed(_native.pack_bits(bits)[None], n, offset)
    if status[0] != 0:
        raise InvalidCodeError("rows are not independent")
    result = _native.unpack_bits(out[0], n).astype(h.dtype)
    h[...] = result
    return np.mod(h, 2), swaps[0]


def _standard_form_gpu(h_1, h_2):
    """CSS condition and both normalisations (css_code.py:47-61) in one ``qcss_css_standard_form`` call."""
    status, n_1, n_2, _ = _native.css_standard_form_bits(h_1.astype(np.uint8), h_2.astype(np.uint8))
    if status == _native.FORM_NOT_CSS:
        raise ValueError("C_2 dual code must be a subspace of C_1")
    if status in (_native.FORM_FEW_COLUMNS_C1, _native.FORM_FEW_COLUMNS_C2):
        raise ValueError("not enough columns")
    if status != 0:
        raise InvalidCodeError("rows are not independent")
    return n_1.astype('int'), n_2.astype('int')


def _layer_keys(parity_check, supports):
    """Big-endian keys (bin_matrix.vec_to_int of H.e mod 2) for a block of supports."""
    m = parity_check.shape[0]
    if supports.shape[1] == 0:
        synd = np.zeros((supports.shape[0], m), dtype=np.int64)
    else:
        synd = parity_check.T[supports].sum(axis=1) % 2          # (block, m)
    keys = np.zeros(supports.shape[0], dtype=np.int64)
    for i in range(m):                                            # int64 wrap == reference's
        keys = (keys << 1) + synd[:, i]
    return keys


def syndrome_table(parity_check):
    """Unique-decoding radius t and the table {syndrome key -> minimum-weight error}.

    Mirrors css_code.py:715-735: weight layers w = 0, 1, ... in ``weight_w_vectors`` order; the
    first repeated syndrome (within the layer or against lower layers) ends the search and
    returns ``(w - 1, table)`` without the partial layer.  Keys are ``np.int64`` big-endian
    integers, values fresh dtype='int' vectors, insertion order = (weight, lexicographic
    support), as in the reference.  The per-vector Python loop of the reference is replaced by
    one integer matrix product per block of supports.
    """
    parity_check = np.asarray(parity_check)
    _, n = parity_check.shape
    h = np.mod(parity_check, 2).astype(np.int64)
    table = {}
    seen = set()
    for w in range(n + 1):
        layer_keys, layer_supports = [], []
        layer_seen = set()
        combos = itertools.combinations(range(n), w)
        while True:
            block = list(itertools.islice(combos, _LAYER_CHUNK))
            if not block:
                break
            supports = np.array(block, dtype=np.intp).reshape(len(block), w)
            keys = _layer_keys(h, supports)
            key_list = keys.tolist()
            block_set = set(key_list)
            if (len(block_set) != len(key_list) or not block_set.isdisjoint(seen)
                    or not block_set.isdisjoint(layer_seen)):
                return w - 1, table
            layer_seen |= block_set
            layer_keys.append(keys)
            layer_supports.append(supports)
        for keys, supports in zip(layer_keys, layer_supports):
            for key, support in zip(keys, supports):
                vec = np.zeros(n, dtype='int')
                vec[support] = 1
                table[key] = vec
        seen |= layer_seen
    return n, table


def syndrome_table_gpu(parity_check, max_entries=1 << 28):
    """``syndrome_table`` with the weight layers enumerated and checked for collisions on the GPU
    (``qcss_table_build``): same ``(t, table)``, same ``np.int64`` big-endian keys, same
    insertion order (weight, then lexicographic support), values fresh dtype='int' vectors.  Needs
    n <= 64 and m <= 62 (the reference's keys wrap at 64 bits anyway); raises ``NativeLibraryError``
    without the CUDA library -- there is no silent fallback inside this function."""
    parity_check = np.asarray(parity_check)
    _, n = parity_check.shape
    t, keys, supports = _native.syndrome_table_arrays(np.mod(parity_check, 2).astype(np.uint8), max_entries)
    vecs = ((supports[:, None] >> np.arange(n, dtype=np.uint64)[None, :]) & np.uint64(1)).astype('int')
    return t, dict(zip(keys, vecs))


def codes_equal(parity_check_1, parity_check_2) -> bool:
    """Two parity checks span the same code iff their RREFs agree (css_code.py:838-844).
    Both RREFs run on the GPU in one batched call."""
    parity_check_1 = np.asarray(parity_check_1)
    parity_check_2 = np.asarray(parity_check_2)
    if parity_check_1.shape != parity_check_2.shape:
        return False
    both = np.stack([np.mod(parity_check_1, 2), np.mod(parity_check_2, 2)]).astype(np.uint8)
    out, _, _ = bin_matrix.rref_batched(both)
    return bool(np.array_equal(out[0], out[1]))


def is_doubly_even(mat):
    """True when every row weight is a multiple of 4 (css_code.py:846-850)."""
    return not np.any(np.mod(np.sum(mat, axis=1), 4))


def quil_classical_correct(prog, codeword, errors, scratch, parity_check, syndromes):
    """css_code.quil_classical_correct (css_code.py:649-685): append the classical Quil that extracts
    the syndrome of ``codeword ^ errors`` and folds the matching table correction into ``errors``.
    ``codeword`` / ``errors`` / ``scratch`` are ``quil_text.Chunk`` register slices (or anything
    indexable to ``name[i]`` cells with slicing); instructions are appended to ``prog`` one text
    line at a time, in the reference's order."""
    from . import quil_text
    for line in quil_text.quil_classical_correct(codeword, errors, scratch, parity_check, syndromes):
        prog += [line] if isinstance(prog, list) else line
    return prog


def quil_classical_detect(prog, codeword, errors, outcome, scratch, parity_check):
    """css_code.quil_classical_detect (css_code.py:687-713): ``outcome`` = 1 iff the syndrome of
    ``codeword ^ errors`` is non-zero."""
    from . import quil_text
    for line in quil_text.quil_classical_detect(codeword, errors, outcome, scratch, parity_check):
        prog += [line] if isinstance(prog, list) else line
    return prog


class _DevicePath:
    """The GPU hot path over the attributes every reference ``CSSCode`` carries (css_code.py:63-72,
    124-161): ``n``, ``parity_check_c1/2``, ``_c1/_c2_syndromes``, ``x/z_operator_matrix()``.
    ``CSSCode`` below inherits it; ``attach(obj)`` puts it next to an EXISTING reference object."""

    _device_code = None

    def _side(self, which):
        if which == 2:
            return self.parity_check_c2, self._c2_syndromes
        if which == 1:
            return self.parity_check_c1, self._c1_syndromes
        raise ValueError("which must be 1 (Z errors, C_1) or 2 (X errors, C_2)")

    @property
    def device(self):
        """The device-resident code object (created on first use)."""
        if self._device_code is None:
            self._device_code = _native.DeviceCode.from_csscode(self)
        return self._device_code

    def syndrome_histogram(self, errors, which):
        """Counts of every syndrome key over a batch: ``hist[vec_to_int(H.e mod 2)]`` (uint64[2^m]),
        computed on the device (``qcss_syndrome_hist``).  The keys are those of ``_c1/_c2_syndromes``."""
        errors = np.asarray(errors)
        return self.device.syndrome_hist_planes(_planes.pack_planes(errors), errors.shape[0], which)

    def specialize(self, compiler="nvrtc"):
        """Compile and attach decode kernels specialised for this code (``specialize.py``): the static
        family that runs Steane / QRM-15 / Golay-23 at the HBM roofline, for any code with n <= 32 and
        m <= 16.  In process with NVRTC by default (no toolkit, no subprocess; ``compiler="nvcc"`` runs nvcc);
        one compilation per distinct code, cached on disk.  Returns the kernel family name."""
        from . import specialize as _spec
        return _spec.specialize(self.device, compiler)

    def syndromes(self, errors, which):
        """Batched ``np.mod(np.matmul(parity_check, e), 2)`` (css_code.py:728) for a
        (shots, n) 0/1 array; returns (shots, m) uint8."""
        self._side(which)
        return self.device.syndrome_shots(errors, which)       # transposed to bit planes on the device

    def decode(self, errors, which):
        """Lookup-decode a (shots, n) batch of one Pauli type with the semantics of
        quil_classical_correct (css_code.py:649-685): correction = table.get(key, 0).
        Returns dict(correction (shots, n), flip (shots,), miss (shots,), tally)."""
        corr, flip, miss, tally = self.device.decode_shots(errors, which)
        return dict(correction=corr, flip=flip, miss=miss, tally=tally)

    def decode_xz(self, x_errors, z_errors):
        """Tallies (shots, fail_x, fail_z, fail_any, miss_x, miss_z) for a shared batch."""
        return self.device.decode_xz_shots(x_errors, z_errors)

    def decode_xz_sparse(self, events, shots):
        """The same tallies for a batch given in sparse form: uint64 events ``shot << 18 | qubit << 2 | pauli``
        sorted by shot (``planes.events_from_arrays``); 30x fewer bytes than bit planes at p = 1e-3 and
        event-driven on the device (``qcss_decode_xz_sparse``)."""
        return self.device.decode_xz_sparse(events, shots)

    def monte_carlo(self, p, shots, seed=0, first_shot=0):
        """Depolarising-noise Monte Carlo fully on the device (sampler fused into the decode
        kernel); returns the tally dict."""
        return self.device.mc_run(p, shots, seed, first_shot)

    def error_correct_monte_carlo(self, p_data, p_ancilla, rounds, shots, seed=0, first_shot=0):
        """Pauli-frame Monte Carlo of ``rounds`` repetitions of the Steane error-correction gadget that
        ``error_correct`` emits in the reference (css_code.py:436-470): each round the data block
        suffers depolarising(p_data), a |+>_L and a |0>_L ancilla with depolarising(p_ancilla) extract
        the X and Z syndromes through transversal CNOTs (with their back-action on the data), and the
        Pauli frames are updated as ``quil_classical_correct`` does with ``_c2_syndromes`` /
        ``_c1_syndromes``; the residual after the last round is decoded ideally.  Entirely on the
        device (``qcss_ec_run``: sampler, syndromes, lookups and frames in registers); returns the tally
        dict of ``monte_carlo``.  ``rounds=1, p_ancilla=0`` is ``monte_carlo(p_data)`` bit for bit.
        Model and Philox stream layout: ``csrc/ec_rounds.cuh``."""
        return self.device.ec_run(p_data, p_ancilla, rounds, shots, seed, first_shot)

    def sample_errors(self, p, shots, seed=0, first_shot=0):
        """The error batch ``monte_carlo`` would draw, as ((shots, n), (shots, n)) uint8."""
        ex, ez = self.device.mc_sample(p, shots, seed, first_shot)
        return _planes.unpack_planes(ex, shots), _planes.unpack_planes(ez, shots)


class AttachedCode(_DevicePath):
    """The device path bound to an existing code object -- typically a ``CSSCode`` built by the UNMODIFIED
    reference (``ftqc.rewrite_program`` users keep that object for gate emission, ftqc.py:42).  Anything
    with the reference's attributes works: ``n``, ``parity_check_c1``, ``parity_check_c2``,
    ``_c1_syndromes``, ``_c2_syndromes``, ``x_operator_matrix()``, ``z_operator_matrix()``
    (css_code.py:63-72, 124-161).  Other attribute reads fall through to the wrapped object, so the
    wrapper can stand in for it."""

    def __init__(self, code):
        for name in ("n", "parity_check_c1", "parity_check_c2", "_c1_syndromes", "_c2_syndromes",
                     "x_operator_matrix", "z_operator_matrix"):
            if not hasattr(code, name):
                raise TypeError(f"cannot attach: object has no attribute {name!r} (css_code.py:63-72, 124-161)")
        self._code = code

    def __getattr__(self, name):
        return getattr(self._code, name)


def attach(code):
    """``attach(reference_css_code).monte_carlo(...)``: bind the CUDA hot path to an existing object."""
    return AttachedCode(code)


class CSSCode(_DevicePath):
    """Calderbank-Shor-Steane code built from two classical parity checks (css_code.py:19-75).

    The constructor performs the reference's host-side numerics (validation, standard form,
    syndrome tables); the per-shot work -- syndromes, lookup decode, logical check, Monte-Carlo
    sampling -- runs in CUDA kernels on bit-plane batches through ``libqcss.so``.
    """

    def __init__(self, parity_check_c1, parity_check_c2, table_builder=None, standard_form=None,
                 allow_multi_logical=False):
        """``table_builder`` (extension, not in the reference): ``"gpu"`` builds both syndrome tables
        with the device weight-layer search (``syndrome_table_gpu``); the default is the host search
        (``syndrome_table``), which needs no GPU.  Both give identical tables.
        ``standard_form`` (extension): ``"gpu"`` runs the CSS condition and both normalisations with
        their qubit swaps on the device (``qcss_css_standard_form``); default host numpy.  Identical
        matrices and exceptions either way.
        ``allow_multi_logical`` (extension): accept k = n - r_1 - r_2 > 1 where the reference raises
        ``InvalidCodeError`` (css_code.py:74-75).  ``x/z_operator_matrix`` then have k rows and the device path counts
        a shot as failed when ANY logical operator flips."""
        r_1, n_1 = parity_check_c1.shape
        r_2, n_2 = parity_check_c2.shape
        if n_1 != n_2:
            raise ValueError("C_1 and C_2 must have the same code word length")
        if table_builder not in (None, "host", "gpu"):
            raise ValueError("table_builder must be None, 'host' or 'gpu'")
        if standard_form not in (None, "host", "gpu"):
            raise ValueError("standard_form must be None, 'host' or 'gpu'")
        build_table = syndrome_table_gpu if table_builder == "gpu" else syndrome_table

        h_1 = np.mod(np.array(parity_check_c1, dtype='int'), 2)
        h_2 = np.mod(np.array(parity_check_c2, dtype='int'), 2)
        if not np.array_equal(h_1, parity_check_c1):
            raise ValueError("C_1 parity check matrix must be binary")
        if not np.array_equal(h_2, parity_check_c2):
            raise ValueError("C_2 parity check matrix must be binary")

        if standard_form == "gpu":
            h_1, h_2 = _standard_form_gpu(h_1, h_2)
        else:
            # CSS condition: every C_2 check is orthogonal to every C_1 check (css_code.py:47-49).
            if np.any(np.mod(h_1 @ h_2.T, 2)):
                raise ValueError("C_2 dual code must be a subspace of C_1")

            # Standard form H_1 = [I A_1 A_2], H_2 = [D I E]; qubit swaps found while normalising one
            # matrix are replayed on the other (css_code.py:55-61).
            h_1, swaps = normalize_parity_check(h_1, offset=0)
            for pair in swaps:
                swap_columns(h_2, pair)
            h_2, swaps = normalize_parity_check(h_2, offset=r_1)
            for pair in swaps:
                swap_columns(h_1, pair)

        self._n = n_1
        self._k = n_1 - r_1 - r_2
        self.r_1 = r_1
        self.r_2 = r_2
        self.parity_check_c1 = h_1
        self.parity_check_c2 = h_2
        t_1, self._c1_syndromes = build_table(h_1)
        t_2, self._c2_syndromes = build_table(h_2)
        self._t = min(t_1, t_2)
        self._transversal_cache = None
        self._device_code = None

        if self.k != 1 and not (allow_multi_logical and self.k > 1):
            raise InvalidCodeError("currently only supports CSS codes for a single logical qubit")

    # ---- reference surface ---------------------------------------------------------------
    @property
    def n(self):
        """Number of physical qubits per code block."""
        return self._n

    @property
    def k(self):
        """Number of logical qubits per code block."""
        return self._k

    @property
    def t(self):
        """Maximum number of errors per code block that can be corrected."""
        return self._t

    @property
    def _transversal_gates(self):
        # css_code.py:182-201.  Evaluated on first use rather than in __init__ because
        # codes_equal runs its two RREFs on the GPU; the value is the reference's.
        if self._transversal_cache is None:
            self._transversal_cache = self._determine_transversal_gates(self.parity_check_c1, self.parity_check_c2)
        return self._transversal_cache

    def _determine_transversal_gates(self, parity_check_c1, parity_check_c2):
        """css_code.py:182-201: I and CNOT always; H and CZ when C_1 = C_2; S when that code is doubly even.
        A frozenset, as in the reference."""
        gates = ['I', 'CNOT']
        if codes_equal(parity_check_c1, parity_check_c2):
            gates += ['H', 'CZ']
            if is_doubly_even(parity_check_c1):
                gates.append('S')
        return frozenset(gates)

    def is_transversal(self, gate_name: str) -> bool:
        """Whether the gate is fault tolerant when applied qubit by qubit (css_code.py:174-180)."""
        return gate_name in self._transversal_gates

    def z_operator_matrix(self):
        """Check matrix [A_2^T 0 I] of the logical Z operators (css_code.py:124-136)."""
        n, r_1, r_2, k = self.n, self.r_1, self.r_2, self.k
        check_mat = np.zeros((k, n), dtype='int')
        check_mat[:, :r_1] = self.parity_check_c1[:, r_1 + r_2:].T
        check_mat[:, r_1 + r_2:] = np.identity(k)
        return check_mat

    def x_operator_matrix(self):
        """Check matrix [0 E^T I] of the logical X operators (css_code.py:149-161)."""
        n, r_1, r_2, k = self.n, self.r_1, self.r_2, self.k
        check_mat = np.zeros((k, n), dtype='int')
        check_mat[:, r_1:r_1 + r_2] = self.parity_check_c2[:, r_1 + r_2:].T
        check_mat[:, r_1 + r_2:] = np.identity(k)
        return check_mat

    @property
    def encode_scratch_size(self) -> int:
        return 2 * self.n - max(self.r_1, self.r_2) + 4          # css_code.py:595-597

    @property
    def measure_scratch_size(self) -> int:
        return self.encode_scratch_size + 2 * self.t + 1          # css_code.py:591-593

    @property
    def error_correct_scratch_size(self) -> int:
        return self.encode_scratch_size                           # css_code.py:539-540


class SyndromeCode:
    """A pair of check matrices used for syndrome extraction only (no tables, no logical rows).

    For codes the reference constructor rejects (k != 1 or dependent rows, css_code.py:74-75,
    825-826), e.g. hypergraph-product codes: the only reference operation that applies is the
    raw syndrome idiom css_code.py:728.  ``which`` = 1 selects parity_check_c1, 2 selects
    parity_check_c2, as for CSSCode."""

    def __init__(self, parity_check_c1, parity_check_c2):
        h_1 = np.mod(np.array(parity_check_c1, dtype='int'), 2)
        h_2 = np.mod(np.array(parity_check_c2, dtype='int'), 2)
        if h_1.shape[1] != h_2.shape[1]:
            raise ValueError("C_1 and C_2 must have the same code word length")
        self.n = h_1.shape[1]
        self.parity_check_c1, self.parity_check_c2 = h_1, h_2
        self._device_code = None

    @property
    def device(self):
        if self._device_code is None:
            self._device_code = _native.DeviceCode(self.n, self.parity_check_c1,
                                                   self.parity_check_c2, None, None, None, None)
        return self._device_code

    def syndromes(self, errors, which):
        return self.device.syndrome_shots(errors, which)

    def sample_syndromes(self, p, shots, seed=0, first_shot=0, return_errors=False):
        """Depolarising(p) errors drawn on the device and turned into syndromes in the same kernel
        (``qcss_sample_syndrome_tiles``: the sampled errors never leave shared memory).  Returns
        ``(s_x, s_z)`` -- (shots, m2) syndromes of the X errors under parity_check_c2 and (shots, m1) of the
        Z errors under parity_check_c1 -- and with ``return_errors`` also ``(e_x, e_z)`` as (shots, n)."""
        out = self.device.sample_syndrome_tiles(p, shots, seed, first_shot, errors=return_errors)
        res = tuple(_planes.unpack_tiles(t, shots) for t in out)
        return res

    def syndromes_tiled(self, errors, which):
        """The same result through the tile-major layout (``planes.pack_tiles`` -> ``qcss_syndrome_tiles``):
        the layout the sparse any-size kernel streams fastest (one bulk copy per part-tile)."""
        errors = np.asarray(errors)
        shots = errors.shape[0]
        s_tiles = self.device.syndrome_tiles(_planes.pack_tiles(errors), shots, which)
        return _planes.unpack_tiles(s_tiles, shots)
