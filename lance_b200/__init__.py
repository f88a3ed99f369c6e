"""lance_b200 -- B200-native IVF-PQ hot path behind lancedb/lance's own operator API.

Host-side mirror (Python, test/bench tooling) of the reference's
`lance-index::vector::{kmeans,ivf,pq,flat}` + `lance-linalg::distance` functions: same names,
argument meaning and error behaviour; every call goes straight through the C ABI
(include/lance_b200.h) into hand-written sm_100a kernels.  No CPU fallback exists.

Inputs may be numpy arrays (host memory) or `DeviceArray`s (resident in HBM).
"""
import ctypes as C

import numpy as np

from ._lib import (BF16, COSINE, DOT, F16, F32, L2, METRICS, U8, BuildParams, BuildStats, FlatBuildParams,
                   DeviceArray, KMeansParams as _CKMeansParams, LanceB200Error, PinnedArray,
                   PQParams as _CPQParams, as_ptr, check, device_count, lib)

__all__ = ["device_count", "DeviceArray", "PinnedArray", "LanceB200Error", "train_kmeans",
           "compute_partitions", "kmeans_find_partitions", "compute_residual", "normalize_fsl",
           "l2_distance_batch", "dot_distance_batch", "cosine_distance_batch", "PQBuildParams", "ProductQuantizer",
           "build_distance_table_l2", "compute_pq_distance", "flat_topk", "IvfPqIndex",
           "IvfBuildParams", "IvfFlatIndex", "launch_count", "profile"]


def _metric(m):
    if isinstance(m, str):
        return METRICS[m.lower()]
    return int(m)


def _shape2(a):
    return a.shape


def _f32(a):
    if isinstance(a, (DeviceArray, PinnedArray)):
        assert a.dtype == np.float32
        return a
    return np.ascontiguousarray(a, dtype=np.float32)


_DTYPES = {np.dtype(np.float32): F32, np.dtype(np.float16): F16, np.dtype(np.uint8): U8}


def _typed(a, bf16=False):
    """(array, lb2_dtype): f32 / f16 / u8 buffers are passed as they are (bf16 = uint16 + bf16=True)."""
    if bf16:
        if isinstance(a, (DeviceArray, PinnedArray)):
            assert np.dtype(a.dtype) == np.uint16
            return a, BF16
        return np.ascontiguousarray(a, dtype=np.uint16), BF16
    if isinstance(a, (DeviceArray, PinnedArray)):
        return a, _DTYPES[np.dtype(a.dtype)]
    a = np.asarray(a)
    if a.dtype not in _DTYPES:
        a = a.astype(np.float32)
    return np.ascontiguousarray(a), _DTYPES[a.dtype]


def _model_np(dt):
    return {F32: np.float32, F16: np.float16, BF16: np.uint16, U8: np.float32}[dt]


def set_device(i):
    check(lib().lb2_set_device(C.c_int(i)))


def synchronize():
    check(lib().lb2_synchronize())


def trim_memory():
    """lb2_trim_memory: release this thread's cached bulk-copy staging buffer and the pool's free blocks."""
    check(lib().lb2_trim_memory())


def set_stream(cuda_stream=None):
    """lb2_set_stream: order every later call of this thread on a caller-owned cudaStream_t (int handle); None =
    back to the thread's private stream."""
    check(lib().lb2_set_stream(C.c_void_p(cuda_stream)))


def launch_count(reset=False):
    n = C.c_uint64(0)
    check(lib().lb2_launch_count(C.byref(n), C.c_int(int(reset))))
    return n.value


class profile:
    """Per-kernel-family CUDA-event timing (lb2_profile_*)."""

    @staticmethod
    def enable(on=True):
        check(lib().lb2_profile_enable(C.c_int(int(on))))

    @staticmethod
    def reset():
        check(lib().lb2_profile_reset())

    @staticmethod
    def dump():
        """{name: (launches, total_ms)} for every kernel family seen since the last reset"""
        n = lib().lb2_profile_dump(None, 0)
        buf = C.create_string_buffer(n + 1)
        lib().lb2_profile_dump(buf, n + 1)
        out = {}
        for line in buf.value.decode().splitlines():
            name, cnt, ms = line.split("\t")
            out[name] = (int(cnt), float(ms))
        return out

    @staticmethod
    def get(name):
        n, ms = C.c_uint64(0), C.c_double(0)
        check(lib().lb2_profile_get(name.encode(), C.byref(n), C.byref(ms)))
        return n.value, ms.value


def timer_start():
    check(lib().lb2_timer_start())


def timer_stop():
    ms = C.c_float(0)
    check(lib().lb2_timer_stop(C.byref(ms)))
    return ms.value


# ---- lance-linalg ---------------------------------------------------------------------------
def l2_distance_batch(frm, to, dimension):
    """lance_linalg::distance::l2_distance_batch (l2.rs:194-203)."""
    return _distance_batch(frm, to, dimension, L2)


def dot_distance_batch(frm, to, dimension):
    """lance_linalg::distance::dot_distance_batch (dot.rs:164-172): 1 - dot."""
    return _distance_batch(frm, to, dimension, DOT)


def cosine_distance_batch(frm, to, dimension):
    """lance_linalg::distance::cosine_distance_batch (cosine.rs:266-290)."""
    return _distance_batch(frm, to, dimension, COSINE)


def _distance_batch(frm, to, d, metric):
    """f32 / f16 / u8 inputs keep their element type (u8 L2 = the reference's integer sum, l2.rs:44-49)."""
    to, dt = _typed(to)
    frm = np.ascontiguousarray(frm, dtype=to.dtype) if not isinstance(frm, (DeviceArray, PinnedArray)) else frm
    n = int(np.prod(to.shape)) // d
    out = np.empty(n, np.float32)
    fp, _k1 = as_ptr(frm)
    tp, _k2 = as_ptr(to)
    check(lib().lb2_distance_batch(fp, tp, C.c_uint64(n), C.c_uint32(d), C.c_int(dt),
                                   C.c_int(metric), C.c_void_p(out.ctypes.data)))
    return out


def normalize_fsl(vectors):
    """lance_linalg::kernels::normalize_fsl (kernels.rs:201-211)."""
    vectors = _f32(vectors)
    n, d = vectors.shape
    out = np.empty((n, d), np.float32)
    vp, _k = as_ptr(vectors)
    check(lib().lb2_normalize(vp, C.c_uint64(n), C.c_uint32(d), C.c_int(F32),
                              C.c_void_p(out.ctypes.data)))
    return out


# ---- lance-index::vector::kmeans -------------------------------------------------------------
class KMeans:
    """Result of train_kmeans (kmeans.rs:527-545: centroids, dimension, distance_type, loss)."""

    def __init__(self, centroids, dimension, distance_type, loss, iters):
        self.centroids, self.dimension, self.distance_type = centroids, dimension, distance_type
        self.loss, self.iters = loss, iters


def train_kmeans(array, dimension, k, max_iters=50, redos=1, distance_type="l2", sample_rate=256,
                 balance_factor=0.0, tolerance=1e-4, seed=0, centroids=None):
    """lance_index::vector::kmeans::train_kmeans (kmeans.rs:1309-1347)."""
    array, dt = _typed(array)
    n = int(np.prod(array.shape)) // dimension
    p = _CKMeansParams()
    lib().lb2_kmeans_params_default(C.byref(p))
    p.max_iters, p.redos, p.sample_rate, p.seed = max_iters, redos, sample_rate, seed
    p.balance_factor, p.tolerance, p.metric = balance_factor, tolerance, _metric(distance_type)
    init = None if centroids is None else np.ascontiguousarray(centroids, dtype=_model_np(dt))
    ip, _k0 = as_ptr(init)
    p.init_centroids = ip.value if ip is not None else None
    out = np.empty((k, dimension), _model_np(dt))
    loss, iters = C.c_double(0), C.c_uint32(0)
    ap, _k1 = as_ptr(array)
    check(lib().lb2_kmeans_train(ap, C.c_uint64(n), C.c_uint32(dimension), C.c_int(dt),
                                 C.c_uint32(k), C.byref(p), C.c_void_p(out.ctypes.data),
                                 C.byref(loss), C.byref(iters)))
    return KMeans(out, dimension, distance_type, loss.value, iters.value)


def compute_partitions(centroids, vectors, distance_type="l2", bf16=False):
    """compute_partitions_arrow_array (kmeans.rs:1187-1246) ->
    (part_ids u32[n], dists f32[n], valid bool[n]); valid False == the reference's None.
    The model has the rows' element type (bf16=True: both are uint16 bit patterns)."""
    vectors, dt = _typed(vectors, bf16)
    centroids = np.ascontiguousarray(centroids, dtype=_model_np(dt))
    k, d = centroids.shape
    n = vectors.shape[0]
    part = np.empty(n, np.uint32)
    dist = np.empty(n, np.float32)
    valid = np.empty(n, np.uint8)
    cp, _k1 = as_ptr(centroids)
    vp, _k2 = as_ptr(vectors)
    check(lib().lb2_compute_partitions(cp, C.c_uint32(k), C.c_uint32(d), C.c_int(dt),
                                       C.c_int(_metric(distance_type)), vp, C.c_uint64(n),
                                       C.c_void_p(part.ctypes.data), C.c_void_p(dist.ctypes.data),
                                       C.c_void_p(valid.ctypes.data)))
    return part, dist, valid.astype(bool)


def kmeans_find_partitions(centroids, queries, nprobes, distance_type="l2"):
    """kmeans_find_partitions_arrow_array (kmeans.rs:1076-1158), batched over queries."""
    centroids, queries = _f32(centroids), _f32(queries)
    single = queries.ndim == 1
    if single:
        queries = queries.reshape(1, -1)
    k, d = centroids.shape
    nq = queries.shape[0]
    ids = np.empty((nq, nprobes), np.uint32)
    dists = np.empty((nq, nprobes), np.float32)
    cp, _k1 = as_ptr(centroids)
    qp, _k2 = as_ptr(queries)
    check(lib().lb2_find_partitions(cp, C.c_uint32(k), C.c_uint32(d), C.c_int(F32),
                                    C.c_int(_metric(distance_type)), qp, C.c_uint64(nq),
                                    C.c_uint32(nprobes), C.c_void_p(ids.ctypes.data),
                                    C.c_void_p(dists.ctypes.data)))
    return (ids[0], dists[0]) if single else (ids, dists)


def compute_residual(centroids, vectors, partitions):
    """lance_index::vector::residual::compute_residual (residual.rs:111-154)."""
    centroids, vectors = _f32(centroids), _f32(vectors)
    k, d = centroids.shape
    n = vectors.shape[0]
    parts = np.ascontiguousarray(partitions, dtype=np.uint32)
    out = np.empty((n, d), np.float32)
    cp, _k1 = as_ptr(centroids)
    vp, _k2 = as_ptr(vectors)
    check(lib().lb2_compute_residual(cp, C.c_uint32(k), C.c_uint32(d), C.c_int(F32), vp,
                                     C.c_uint64(n), C.c_void_p(parts.ctypes.data),
                                     C.c_void_p(out.ctypes.data)))
    return out


# ---- lance-index::vector::pq -----------------------------------------------------------------
class PQBuildParams:
    """lance_index::vector::pq::builder::PQBuildParams (pq/builder.rs:27-59)."""

    def __init__(self, num_sub_vectors=16, num_bits=8, max_iters=50, kmeans_redos=1, codebook=None,
                 sample_rate=256, seed=0):
        self.num_sub_vectors, self.num_bits, self.max_iters = num_sub_vectors, num_bits, max_iters
        self.kmeans_redos, self.codebook, self.sample_rate, self.seed = kmeans_redos, codebook, sample_rate, seed

    def _c(self):
        p = _CPQParams()
        lib().lb2_pq_params_default(C.byref(p))
        p.num_sub_vectors, p.num_bits, p.max_iters = self.num_sub_vectors, self.num_bits, self.max_iters
        p.kmeans_redos, p.sample_rate, p.seed = self.kmeans_redos, self.sample_rate, self.seed
        self._cb = None if self.codebook is None else _f32(self.codebook)
        cp, _ = as_ptr(self._cb)
        p.codebook = cp.value if cp is not None else None
        return p

    def build(self, data, distance_type="l2"):
        """PQBuildParams::build (pq/builder.rs:162-194) -> ProductQuantizer."""
        data = _f32(data)
        n, d = data.shape
        p = self._c()
        out = np.empty((self.num_sub_vectors, 1 << self.num_bits, d // self.num_sub_vectors), np.float32)
        iters = np.zeros(self.num_sub_vectors, np.uint32)
        dp, _k = as_ptr(data)
        check(lib().lb2_pq_train(dp, C.c_uint64(n), C.c_uint32(d), C.c_int(F32),
                                 C.c_int(_metric(distance_type)), C.byref(p),
                                 C.c_void_p(out.ctypes.data), C.c_void_p(iters.ctypes.data)))
        pq = ProductQuantizer(self.num_sub_vectors, self.num_bits, d, out, distance_type)
        pq.train_iters = iters
        return pq


class ProductQuantizer:
    """lance_index::vector::pq::ProductQuantizer (pq.rs:42-48); codebook [M][2^nbits][d/M]."""

    def __init__(self, num_sub_vectors, num_bits, dimension, codebook, distance_type="l2"):
        self.num_sub_vectors, self.num_bits, self.dimension = num_sub_vectors, num_bits, dimension
        self.codebook = _f32(codebook).reshape(num_sub_vectors, 1 << num_bits, dimension // num_sub_vectors)
        self.distance_type = distance_type

    def quantize(self, vectors, centroids=None, part_ids=None):
        """Quantization::quantize (pq.rs:430) = transform_impl (pq.rs:116-191); with
        centroids+part_ids the residual transform (residual.rs:161-205) is fused in."""
        vectors = _f32(vectors)
        n, d = vectors.shape
        out = np.empty((n, self.num_sub_vectors // 2 if self.num_bits == 4 else self.num_sub_vectors), np.uint8)
        cent = None if centroids is None else _f32(centroids)
        parts = None if part_ids is None else np.ascontiguousarray(part_ids, dtype=np.uint32)
        vp, _k1 = as_ptr(vectors)
        cp, _k2 = as_ptr(cent)
        pp, _k3 = as_ptr(parts)
        check(lib().lb2_pq_encode(C.c_void_p(self.codebook.ctypes.data), C.c_uint32(self.num_sub_vectors),
                                  C.c_uint32(self.num_bits), C.c_uint32(d), C.c_int(F32),
                                  C.c_int(_metric(self.distance_type)), cp,
                                  C.c_uint32(0 if cent is None else cent.shape[0]), pp, vp, C.c_uint64(n),
                                  C.c_void_p(out.ctypes.data)))
        return out


def build_distance_table_l2(codebook, num_bits, num_sub_vectors, query, distance_type="l2"):
    """lance_index::vector::pq::distance::build_distance_table_l2 / _dot (pq/distance.rs:24-92)."""
    codebook, query = _f32(codebook), _f32(query)
    d = query.size
    out = np.empty(num_sub_vectors << num_bits, np.float32)
    check(lib().lb2_pq_build_lut(C.c_void_p(codebook.ctypes.data), C.c_uint32(num_sub_vectors),
                                 C.c_uint32(num_bits), C.c_uint32(d), C.c_int(_metric(distance_type)),
                                 C.c_void_p(query.ctypes.data), C.c_void_p(out.ctypes.data)))
    return out


def compute_pq_distance(distance_table, num_bits, num_sub_vectors, code_transposed, distance_type="l2"):
    """pq/distance.rs:109-144 on transposed codes [M][n]."""
    lut = _f32(distance_table)
    code = np.ascontiguousarray(code_transposed, dtype=np.uint8)
    n = code.size // num_sub_vectors
    out = np.empty(n, np.float32)
    check(lib().lb2_pq_scan(C.c_void_p(lut.ctypes.data), C.c_uint32(num_sub_vectors),
                            C.c_uint32(num_bits), C.c_int(_metric(distance_type)),
                            C.c_void_p(code.ctypes.data), C.c_uint64(n), C.c_void_p(out.ctypes.data)))
    return out


def compute_pq_distance_4bit(distance_table, num_sub_vectors, code_transposed, k_hint, distance_type="l2"):
    """pq/distance.rs:147-242 on transposed packed codes [M/2][n]; distance_table [M][16]."""
    lut = _f32(distance_table)
    code = np.ascontiguousarray(code_transposed, dtype=np.uint8)
    n = code.size // (num_sub_vectors // 2)
    out = np.empty(n, np.float32)
    check(lib().lb2_pq_scan_4bit(C.c_void_p(lut.ctypes.data), C.c_uint32(num_sub_vectors),
                                 C.c_int(_metric(distance_type)), C.c_void_p(code.ctypes.data), C.c_uint64(n),
                                 C.c_uint64(k_hint), C.c_void_p(out.ctypes.data)))
    return out


def flat_topk(dists, row_ids, k, lower_bound=None, upper_bound=None):
    """FlatIndex::search fast path (flat/index.rs:97-127) over a distance array, optionally with the
    range branch (:100-115)."""
    dists = _f32(dists)
    n = dists.size
    rid = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.uint64)
    oi, od, cnt = np.empty(k, np.uint64), np.empty(k, np.float32), np.zeros(1, np.uint32)
    rp, _k = as_ptr(rid)
    check(lib().lb2_flat_topk_range(C.c_void_p(dists.ctypes.data), rp, C.c_uint64(n), C.c_uint32(k),
                                    C.c_int(lower_bound is not None), C.c_float(lower_bound or 0.0),
                                    C.c_int(upper_bound is not None), C.c_float(upper_bound or 0.0),
                                    C.c_void_p(oi.ctypes.data), C.c_void_p(od.ctypes.data),
                                    C.c_void_p(cnt.ctypes.data)))
    return oi[:cnt[0]], od[:cnt[0]]


def ivfpq_transform(centroids, codebook, vectors, distance_type="l2", num_bits=8):
    """IvfTransformer::transform for IVF_PQ (lance-index/src/vector/ivf.rs:188-236,357)."""
    centroids, codebook, vectors = _f32(centroids), _f32(codebook), _f32(vectors)
    k, d = centroids.shape
    M = codebook.shape[0]
    n = vectors.shape[0]
    part, codes, valid = np.empty(n, np.uint32), np.empty((n, M // 2 if num_bits == 4 else M), np.uint8), np.empty(n, np.uint8)
    cp, _k1 = as_ptr(centroids)
    bp, _k2 = as_ptr(codebook)
    vp, _k3 = as_ptr(vectors)
    check(lib().lb2_ivfpq_transform(cp, C.c_uint32(k), bp, C.c_uint32(M), C.c_uint32(num_bits),
                                    C.c_uint32(d), C.c_int(F32), C.c_int(_metric(distance_type)), vp,
                                    C.c_uint64(n), C.c_void_p(part.ctypes.data),
                                    C.c_void_p(codes.ctypes.data), C.c_void_p(valid.ctypes.data)))
    return part, codes, valid.astype(bool)


# ---- the index (rust/lance/src/index/vector/{builder.rs, ivf/v2.rs}) ---------------------------
class IvfBuildParams:
    """lance_index::vector::ivf::builder::IvfBuildParams (ivf/builder.rs:20-78) + PQBuildParams."""

    def __init__(self, num_partitions=256, num_sub_vectors=16, num_bits=8, max_iters=50,
                 sample_rate=256, pq_max_iters=50, pq_sample_rate=256, seed=0, centroids=None,
                 codebook=None):
        self.num_partitions, self.num_sub_vectors, self.num_bits = num_partitions, num_sub_vectors, num_bits
        self.max_iters, self.sample_rate, self.pq_max_iters = max_iters, sample_rate, pq_max_iters
        self.pq_sample_rate, self.seed, self.centroids, self.codebook = pq_sample_rate, seed, centroids, codebook


class IvfPqIndex:
    """Device-resident IVFIndex<FlatIndex, ProductQuantizer> (ivf/v2.rs:104)."""

    def __init__(self, handle, stats=None):
        self._h = handle
        self.stats = stats

    @classmethod
    def build(cls, data, distance_type="l2", params=None, row_ids=None):
        """IvfIndexBuilder::build (builder.rs:236): create_index("IVF_PQ")."""
        params = params or IvfBuildParams()
        data, dt = _typed(data)
        n, d = data.shape
        bp = BuildParams()
        lib().lb2_ivfpq_build_params_default(C.byref(bp))
        bp.num_partitions = params.num_partitions
        bp.ivf.max_iters, bp.ivf.sample_rate, bp.ivf.seed = params.max_iters, params.sample_rate, params.seed
        bp.pq.num_sub_vectors, bp.pq.num_bits = params.num_sub_vectors, params.num_bits
        bp.pq.max_iters, bp.pq.sample_rate, bp.pq.seed = params.pq_max_iters, params.pq_sample_rate, params.seed + 1000
        bp.seed = params.seed
        keep = []
        if params.centroids is not None:
            c = _f32(params.centroids)
            keep.append(c)
            bp.ivf.init_centroids = as_ptr(c)[0].value
        if params.codebook is not None:
            c = _f32(params.codebook)
            keep.append(c)
            bp.pq.codebook = as_ptr(c)[0].value
        rid = None if row_ids is None else (row_ids if isinstance(row_ids, DeviceArray)
                                            else np.ascontiguousarray(row_ids, dtype=np.uint64))
        h = C.c_void_p()
        st = BuildStats()
        dp, _k1 = as_ptr(data)
        rp, _k2 = as_ptr(rid)
        check(lib().lb2_ivfpq_build(dp, C.c_uint64(n), C.c_uint32(d), C.c_int(dt),
                                    C.c_int(_metric(distance_type)), C.byref(bp), rp, C.byref(h),
                                    C.byref(st)))
        ix = cls(h, st)
        ix._dt = dt
        return ix

    @classmethod
    def from_parts(cls, centroids, codebook, part_ids, codes, row_ids=None, distance_type="l2", num_bits=8):
        """Open an index from a trained model + the shuffle output (user-supplied
        ivf_centroids / pq_codebook / precomputed buffers, ivf/builder.rs:20-60)."""
        centroids, codebook = _f32(centroids), _f32(codebook)
        k, d = centroids.shape
        M = codebook.shape[0]
        h = C.c_void_p()
        check(lib().lb2_index_create(C.c_void_p(centroids.ctypes.data), C.c_uint32(k), C.c_uint32(d),
                                     C.c_int(F32), C.c_int(_metric(distance_type)),
                                     C.c_void_p(codebook.ctypes.data), C.c_uint32(M),
                                     C.c_uint32(num_bits), C.byref(h)))
        ix = cls(h)
        part_ids = np.ascontiguousarray(part_ids, dtype=np.uint32)
        codes = np.ascontiguousarray(codes, dtype=np.uint8)
        rid = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.uint64)
        rp, _k = as_ptr(rid)
        check(lib().lb2_index_load(h, C.c_void_p(part_ids.ctypes.data), C.c_void_p(codes.ctypes.data),
                                   rp, C.c_uint64(part_ids.size)))
        return ix

    def info(self):
        k, d, M, nb, n = C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint32(), C.c_uint64()
        check(lib().lb2_index_info(self._h, C.byref(k), C.byref(d), C.byref(M), C.byref(nb), C.byref(n)))
        return dict(num_partitions=k.value, dimension=d.value, num_sub_vectors=M.value,
                    num_bits=nb.value, num_rows=n.value)

    def export(self, out=None):
        """lb2_index_export; `out` = dict of preallocated arrays (same keys) to write into."""
        i = self.info()
        K, d, M, n = i["num_partitions"], i["dimension"], i["num_sub_vectors"], i["num_rows"]
        if out is not None:
            cent, cb, off, codes, rid = (out[k] for k in ("centroids", "codebook", "part_offsets", "codes", "row_ids"))
        else:
            cent = np.empty((K, d), np.float32)
            nbits = i["num_bits"]
            cb = np.empty((M, 1 << nbits, d // M), np.float32)
            off = np.empty(K + 1, np.uint64)
            codes = np.empty((n, M // 2 if nbits == 4 else M), np.uint8)   # 4-bit: two codes per byte
            rid = np.empty(n, np.uint64)
        check(lib().lb2_index_export(self._h, C.c_void_p(cent.ctypes.data), C.c_void_p(cb.ctypes.data),
                                     C.c_void_p(off.ctypes.data), C.c_void_p(codes.ctypes.data),
                                     C.c_void_p(rid.ctypes.data)))
        return dict(centroids=cent, codebook=cb, part_offsets=off, codes=codes, row_ids=rid)

    def export_partition_transposed(self, partition):
        """lb2_index_export_partition: (codes [code bytes][n_p] column-major, row_ids [n_p]) -- the reference's
        storage / index-file layout of one partition (pq/storage.rs:430-450)."""
        i = self.info()
        cw = i["num_sub_vectors"] // 2 if i["num_bits"] == 4 else i["num_sub_vectors"]
        n = C.c_uint64(0)
        check(lib().lb2_index_export_partition(self._h, C.c_uint32(partition), None, None, C.byref(n)))
        codes, rid = np.empty((cw, n.value), np.uint8), np.empty(n.value, np.uint64)
        check(lib().lb2_index_export_partition(self._h, C.c_uint32(partition), C.c_void_p(codes.ctypes.data),
                                               C.c_void_p(rid.ctypes.data), C.byref(n)))
        return codes, rid

    def search(self, queries, k=10, nprobes=1, out=None):
        """to_table(nearest={q,k,nprobes}) for a batch of queries: IVFIndex::find_partitions +
        search_in_partition + global merge (v2.rs:455-500, scanner.rs:3450-3466).
        `out` = optional (row_ids, dists) arrays (numpy/Pinned/Device) to write into."""
        dt = getattr(self, "_dt", F32)
        if isinstance(queries, (DeviceArray, PinnedArray)):
            assert np.dtype(queries.dtype).itemsize == {F32: 4, F16: 2, BF16: 2, U8: 1}[dt]
        else:
            queries = np.ascontiguousarray(queries, dtype={F32: np.float32, F16: np.float16, U8: np.uint8, BF16: np.uint16}[dt])
        nq = queries.shape[0]
        if out is None:
            ids, dists = np.empty((nq, k), np.uint64), np.empty((nq, k), np.float32)
        else:
            ids, dists = out
        qp, _k1 = as_ptr(queries)
        ip, _k2 = as_ptr(ids)
        dp, _k3 = as_ptr(dists)
        check(lib().lb2_index_search(self._h, qp, C.c_uint64(nq), C.c_uint32(k), C.c_uint32(nprobes),
                                     ip, dp, None))
        return ids, dists

    def search_refine(self, vectors, queries, k=10, nprobes=1, refine_factor=1, out=None):
        """to_table(nearest={..., refine_factor}): PQ candidates re-ranked with exact distances from
        the raw column (scanner.rs:2884-2905).  `vectors` = the indexed column (row id = row number)."""
        dt = getattr(self, "_dt", F32)
        npdt = {F32: np.float32, F16: np.float16, U8: np.uint8, BF16: np.uint16}[dt]
        if not isinstance(queries, (DeviceArray, PinnedArray)):
            queries = np.ascontiguousarray(queries, dtype=npdt)
        if not isinstance(vectors, (DeviceArray, PinnedArray)):
            vectors = np.ascontiguousarray(vectors, dtype=npdt)
        nq = queries.shape[0]
        if out is None:
            ids, dists = np.empty((nq, k), np.uint64), np.empty((nq, k), np.float32)
        else:
            ids, dists = out
        qp, _k1 = as_ptr(queries)
        vp, _k0 = as_ptr(vectors)
        ip, _k2 = as_ptr(ids)
        dp, _k3 = as_ptr(dists)
        check(lib().lb2_index_search_refine(self._h, vp, C.c_uint64(vectors.shape[0]), qp, C.c_uint64(nq),
                                            C.c_uint32(k), C.c_uint32(nprobes), C.c_uint32(refine_factor),
                                            ip, dp, None))
        return ids, dists

    def row_mask(self, allow_row_ids=None, block_row_ids=None):
        """RowIdMask (lance-core/src/utils/mask.rs:84-93) -> bitmap over storage positions
        (uint64 words), built on the device by lb2_index_row_mask."""
        from ._lib import lib as _l
        n = self.info()["num_rows"]
        bm = np.zeros((n + 63) // 64, np.uint64)
        a = None if allow_row_ids is None else np.ascontiguousarray(np.sort(np.asarray(allow_row_ids, dtype=np.uint64)))
        b = None if block_row_ids is None else np.ascontiguousarray(np.sort(np.asarray(block_row_ids, dtype=np.uint64)))
        check(_l().lb2_index_row_mask(self._h,
                                      C.c_void_p(a.ctypes.data) if a is not None and a.size else None,
                                      C.c_uint64(0 if a is None else a.size), C.c_int(a is not None),
                                      C.c_void_p(b.ctypes.data) if b is not None and b.size else None,
                                      C.c_uint64(0 if b is None else b.size), C.c_int(b is not None),
                                      C.c_void_p(bm.ctypes.data)))
        return bm

    def search_ex(self, queries, k=10, nprobes=1, allow_bitmap=None, refine_factor=0, vectors=None, out=None,
                  lower_bound=None, upper_bound=None):
        """lb2_index_search_ex: prefilter (bitmap from row_mask), range bounds and/or refine in one call."""
        from ._lib import SearchParams
        dt = getattr(self, "_dt", F32)
        npdt = {F32: np.float32, F16: np.float16, U8: np.uint8, BF16: np.uint16}[dt]
        if not isinstance(queries, (DeviceArray, PinnedArray)):
            queries = np.ascontiguousarray(queries, dtype=npdt)
        if vectors is not None and not isinstance(vectors, (DeviceArray, PinnedArray)):
            vectors = np.ascontiguousarray(vectors, dtype=npdt)
        nq = queries.shape[0]
        if out is None:
            ids, dists = np.empty((nq, k), np.uint64), np.empty((nq, k), np.float32)
        else:
            ids, dists = out
        qp, _k1 = as_ptr(queries)
        vp, _k0 = as_ptr(vectors)
        bp, _k4 = as_ptr(None if allow_bitmap is None else (allow_bitmap if isinstance(allow_bitmap, (DeviceArray, PinnedArray)) else np.ascontiguousarray(allow_bitmap, dtype=np.uint64)))
        ip, _k2 = as_ptr(ids)
        dp, _k3 = as_ptr(dists)
        sp = SearchParams(k, nprobes, refine_factor, vp.value if vp is not None else None,
                          0 if vectors is None else vectors.shape[0], bp.value if bp is not None else None,
                          int(lower_bound is not None), int(upper_bound is not None),
                          float(lower_bound or 0.0), float(upper_bound or 0.0))
        check(lib().lb2_index_search_ex(self._h, qp, C.c_uint64(nq), C.byref(sp), ip, dp, None))
        return ids, dists

    def search_async(self, queries, out, k=10, nprobes=1, cuda_stream=None, done_event=None, allow_bitmap=None,
                     lower_bound=None, upper_bound=None):
        """lb2_index_search_async: enqueue a search on `cuda_stream` (cudaStream_t handle as int) and return at once.
        queries / out = (ids, dists) are DeviceArray or PinnedArray and must stay alive until the stream is done."""
        from ._lib import SearchParams
        assert all(isinstance(a, (DeviceArray, PinnedArray)) for a in (queries, out[0], out[1]))
        qp, _k1 = as_ptr(queries)
        ip, _k2 = as_ptr(out[0])
        dp, _k3 = as_ptr(out[1])
        bp, _k4 = as_ptr(allow_bitmap)
        sp = SearchParams(k, nprobes, 0, None, 0, bp.value if bp is not None else None,
                          int(lower_bound is not None), int(upper_bound is not None),
                          float(lower_bound or 0.0), float(upper_bound or 0.0))
        check(lib().lb2_index_search_async(self._h, qp, C.c_uint64(queries.shape[0]), C.byref(sp), ip, dp, None,
                                           C.c_void_p(cuda_stream), C.c_void_p(done_event)))

    def update(self, new_centroids=None, part_map=None, add_part_ids=None, add_codes=None, add_row_ids=None,
               remove_row_ids=None):
        """lb2_index_update: merge AssignOp-style changes (append / remove / re-map partitions) into a NEW index."""
        info = self.info()
        dt = getattr(self, "_dt", F32)
        cent = None if new_centroids is None else _typed(new_centroids, dt == BF16)[0]
        new_k = info["num_partitions"] if cent is None else cent.shape[0]
        pm = None if part_map is None else np.ascontiguousarray(part_map, dtype=np.uint32)
        ap = None if add_part_ids is None else np.ascontiguousarray(add_part_ids, dtype=np.uint32)
        ac = None if add_codes is None else np.ascontiguousarray(add_codes, dtype=np.uint8)
        ar = None if add_row_ids is None else np.ascontiguousarray(add_row_ids, dtype=np.uint64)
        rm = None if remove_row_ids is None else np.sort(np.ascontiguousarray(remove_row_ids, dtype=np.uint64))
        ptrs = [as_ptr(x)[0] for x in (cent, pm, ap, ac, ar, rm)]
        h = C.c_void_p()
        check(lib().lb2_index_update(self._h, ptrs[0], C.c_uint32(new_k), ptrs[1], ptrs[2], ptrs[3], ptrs[4],
                                     C.c_uint64(0 if ap is None else ap.size), ptrs[5],
                                     C.c_uint64(0 if rm is None else rm.size), C.byref(h)))
        out = type(self)(h)
        if hasattr(self, "_dt"):
            out._dt = self._dt
        return out

    def repartition(self):
        """lb2_index_repartition: row-sharded index -> the index of the partitions this rank owns (p % nranks == rank),
        by one device all-to-all; a copy without a communicator."""
        h = C.c_void_p()
        check(lib().lb2_index_repartition(self._h, C.byref(h)))
        out = type(self)(h)
        for a in ("_dt",):
            if hasattr(self, a):
                setattr(out, a, getattr(self, a))
        return out

    def search_sharded(self, queries, k=10, nprobes=1, out=None):
        """lb2_index_search_sharded: this index holds ONE RANK'S rows (global row ids); every rank calls with
        the same queries and gets the global top-k (per-rank lists exchanged + merged in the library)."""
        from ._lib import SearchParams
        dt = getattr(self, "_dt", F32)
        npdt = {F32: np.float32, F16: np.float16, U8: np.uint8, BF16: np.uint16}[dt]
        if not isinstance(queries, (DeviceArray, PinnedArray)):
            queries = np.ascontiguousarray(queries, dtype=npdt)
        nq = queries.shape[0]
        if out is None:
            ids, dists = np.empty((nq, k), np.uint64), np.empty((nq, k), np.float32)
        else:
            ids, dists = out
        qp, _k1 = as_ptr(queries)
        ip, _k2 = as_ptr(ids)
        dp, _k3 = as_ptr(dists)
        sp = SearchParams(k, nprobes, 0, None, 0, None, 0, 0, 0.0, 0.0)
        check(lib().lb2_index_search_sharded(self._h, qp, C.c_uint64(nq), C.byref(sp), ip, dp, None))
        return ids, dists

    def close(self):
        if self._h:
            lib().lb2_index_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class IvfFlatIndex(IvfPqIndex):
    """Device-resident IVFIndex<FlatIndex, FlatQuantizer> (IVF_FLAT): exact distances inside the
    probed partitions (lance-index/src/vector/flat/{index,storage}.rs)."""

    @classmethod
    def build(cls, data, distance_type="l2", num_partitions=256, max_iters=50, sample_rate=256, seed=0,
              centroids=None, row_ids=None, bf16=False):
        """bf16=True: `data` is a uint16 array holding bfloat16 bit patterns (numpy has no bf16 dtype)."""
        data, dt = _typed(data, bf16)
        n, d = data.shape
        bp = FlatBuildParams()
        lib().lb2_ivfflat_build_params_default(C.byref(bp))
        bp.num_partitions = num_partitions
        bp.ivf.max_iters, bp.ivf.sample_rate, bp.ivf.seed, bp.seed = max_iters, sample_rate, seed, seed
        keep = None
        if centroids is not None:
            keep = _f32(centroids)
            bp.ivf.init_centroids = as_ptr(keep)[0].value
        rid = None if row_ids is None else (row_ids if isinstance(row_ids, (DeviceArray, PinnedArray))
                                            else np.ascontiguousarray(row_ids, dtype=np.uint64))
        h = C.c_void_p()
        st = BuildStats()
        dp, _k1 = as_ptr(data)
        rp, _k2 = as_ptr(rid)
        check(lib().lb2_ivfflat_build(dp, C.c_uint64(n), C.c_uint32(d), C.c_int(dt),
                                      C.c_int(_metric(distance_type)), C.byref(bp), rp, C.byref(h), C.byref(st)))
        ix = cls(h, st)
        ix._dt = dt
        return ix

    @classmethod
    def from_parts(cls, centroids, part_ids, vectors, row_ids=None, distance_type="l2"):
        centroids, vectors = _f32(centroids), _f32(vectors)
        k, d = centroids.shape
        h = C.c_void_p()
        check(lib().lb2_index_create_flat(C.c_void_p(centroids.ctypes.data), C.c_uint32(k), C.c_uint32(d),
                                          C.c_int(F32), C.c_int(_metric(distance_type)), C.byref(h)))
        ix = cls(h)
        part_ids = np.ascontiguousarray(part_ids, dtype=np.uint32)
        rid = None if row_ids is None else np.ascontiguousarray(row_ids, dtype=np.uint64)
        rp, _k = as_ptr(rid)
        vp, _k2 = as_ptr(vectors)
        check(lib().lb2_index_load_flat(h, C.c_void_p(part_ids.ctypes.data), vp, rp, C.c_uint64(part_ids.size)))
        return ix

    def export(self):
        i = self.info()
        K, d, n = i["num_partitions"], i["dimension"], i["num_rows"]
        cent = np.empty((K, d), np.float32)
        off = np.empty(K + 1, np.uint64)
        # the stored vectors keep the column's element type (bf16 as uint16 bit patterns; u8 columns are held as f32)
        vec = np.empty((n, d), {F32: np.float32, F16: np.float16, BF16: np.uint16, U8: np.float32}[getattr(self, "_dt", F32)])
        rid = np.empty(n, np.uint64)
        check(lib().lb2_index_export_flat(self._h, C.c_void_p(cent.ctypes.data), C.c_void_p(off.ctypes.data),
                                          C.c_void_p(vec.ctypes.data), C.c_void_p(rid.ctypes.data)))
        return dict(centroids=cent, part_offsets=off, vectors=vec, row_ids=rid)
