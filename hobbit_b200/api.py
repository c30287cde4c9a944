"""ctypes binding of include/hobbit_b200.h.  F arrays are numpy uint64 of shape (n, 2) == the reference's
16-byte ``virgo::fieldElement {real, img}``; digests are uint8 arrays of shape (n, 32) == ``_hash``."""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None

c_sz = ctypes.c_size_t
c_vp = ctypes.c_void_p


class HobbitError(RuntimeError):
    pass


def lib_path():
    # HOBBIT_B200_LIB: development override (kernel variants built side by side); the product library is the in-tree one
    return os.environ.get("HOBBIT_B200_LIB") or os.path.join(_HERE, "libhobbit_b200.so")


def load_library():
    """Load the CUDA library.  Fails loudly when it has not been built (no fallback path exists)."""
    global _LIB
    if _LIB is None:
        p = lib_path()
        if not os.path.exists(p):
            raise HobbitError("%s is missing — run `python -m hobbit_b200.build` (nvcc, sm_100a)" % p)
        L = ctypes.CDLL(p)
        L.hb_last_error.restype = ctypes.c_char_p
        L.hb_launch_count.restype = ctypes.c_uint64
        L.hb_transcript_digest.restype = ctypes.c_uint64
        L.hb_stream.restype = c_vp
        L.hb_tensor_device.restype = c_vp
        L.hb_expander_codeword_len.restype = ctypes.c_longlong
        _LIB = L
    return _LIB


class DevF:
    """A device-resident F array: (raw device pointer, element count).  Accepted wherever a numpy F array is; build it from a
    torch CUDA tensor with DevF.from_torch(t) (t: int64/uint64 tensor of shape (n, 2))."""

    def __init__(self, ptr, n, keep=None):
        self.ptr, self.n, self.keep = int(ptr), int(n), keep

    @classmethod
    def from_torch(cls, t):
        assert t.is_cuda and t.is_contiguous() and t.shape[-1] == 2 and t.element_size() == 8
        return cls(t.data_ptr(), t.numel() // 2, keep=t)

    def __len__(self):
        return self.n


def _F(a):
    if isinstance(a, DevF):
        return a
    a = np.ascontiguousarray(np.asarray(a, dtype=np.uint64))
    return a.reshape(-1, 2)


def _ptr(a):
    if a is None:
        return None
    if isinstance(a, DevF):
        return c_vp(a.ptr)
    if isinstance(a, int):          # raw device pointer (e.g. torch tensor .data_ptr())
        return c_vp(a)
    return a.ctypes.data_as(c_vp)


class Context:
    """One CUDA device + stream.  Mirrors the reference's globals (tensor_row_size, linear_time, BUFFER_SPACE) as
    explicit arguments."""

    def __init__(self, device=0):
        self.lib = load_library()
        h = c_vp()
        rc = self.lib.hb_ctx_create(ctypes.byref(h), int(device))
        if rc:
            raise HobbitError("hb_ctx_create failed (rc=%d): no CUDA device / bad ordinal — there is no CPU fallback" % rc)
        self.h = h

    @classmethod
    def from_handle(cls, handle):
        """Wrap an existing hb_ctx* (e.g. the host mirror's: libhobbit_host.so hobbit_c_backend) without owning it."""
        self = cls.__new__(cls)
        self.lib = load_library()
        self.h = c_vp(handle)
        self.borrowed = True
        return self

    def close(self):
        if getattr(self, "h", None):
            if not getattr(self, "borrowed", False):
                self.lib.hb_ctx_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _ck(self, rc):
        if rc:
            raise HobbitError(self.lib.hb_last_error(self.h).decode())

    # ---- plumbing ----
    def sync(self):
        self._ck(self.lib.hb_sync(self.h))

    def launch_count(self):
        return int(self.lib.hb_launch_count(self.h))

    def transcript_digest(self, reset=False):
        return int(self.lib.hb_transcript_digest(self.h, 1 if reset else 0))

    def ubench_pipes(self):
        out = (ctypes.c_double * 3)()
        self._ck(self.lib.hb_ubench_pipes(self.h, out))
        return {"imad_wide": out[0], "alu": out[1], "mix_1_wide_3_alu": out[2], "unit": "warp-inst/s (chip)"}

    def profile(self, on=True):
        self._ck(self.lib.hb_profile_enable(self.h, 1 if on else 0))

    def profile_report(self):
        import json
        self.lib.hb_profile_report.restype = c_sz
        n = self.lib.hb_profile_report(self.h, None, c_sz(0))
        buf = ctypes.create_string_buffer(n + 16)
        self.lib.hb_profile_report(self.h, buf, c_sz(n + 16))
        return json.loads(buf.value.decode())

    def stream(self):
        return int(self.lib.hb_stream(self.h) or 0)

    def pinned(self, shape, dtype):
        """numpy array backed by page-locked host memory (for e2e timing with real H2D/D2H)."""
        n = int(np.prod(shape)) * np.dtype(dtype).itemsize
        p = c_vp()
        self._ck(self.lib.hb_malloc_pinned(self.h, ctypes.byref(p), c_sz(max(n, 1))))
        buf = (ctypes.c_uint8 * max(n, 1)).from_address(p.value)
        arr = np.frombuffer(buf, dtype=dtype, count=int(np.prod(shape))).reshape(shape)
        return arr

    # ---- field ----
    def binop(self, op, a, b=None):
        a = _F(a)
        b = a if b is None else _F(b)
        c = np.empty_like(a)
        self._ck(self.lib.hb_field_binop(self.h, int(op), _ptr(a), _ptr(b), _ptr(c), c_sz(len(a))))
        return c

    def root_of_unity(self, logn):
        o = np.zeros((1, 2), dtype=np.uint64)
        self.lib.hb_root_of_unity(int(logn), _ptr(o))
        return o

    def mimc(self, x, k):
        x, k, o = _F(x), _F(k), np.zeros((1, 2), dtype=np.uint64)
        self.lib.hb_mimc_hash(_ptr(x), _ptr(k), _ptr(o))
        return o

    # ---- NTT ----
    def fft(self, arr, logn, batch=1):
        a = _F(arr).copy()
        self._ck(self.lib.hb_ntt_batch(self.h, _ptr(a), int(logn), c_sz(batch), c_sz(1 << logn)))
        return a

    # ---- expander ----
    def expander_set(self, n, graphs):
        """graphs: {(which, dep): (L, R, deg, nbr u32[L*deg], w u64[L*deg])} with which 0 = C, 1 = D (reference _C/D)."""
        levels = len(graphs) // 2
        LL = ctypes.c_longlong * max(levels, 1)
        PP = c_vp * max(levels, 1)
        LC, RC, LD, RD, nC, wC, nD, wD = LL(), LL(), LL(), LL(), PP(), PP(), PP(), PP()
        keep = []
        degC, degD = 9, 12
        for d in range(levels):
            L, R, deg, nbr, w = graphs[(0, d)]
            LC[d], RC[d], degC = L, R, deg
            nbr = np.ascontiguousarray(nbr, dtype=np.uint32); w = np.ascontiguousarray(w, dtype=np.uint64)
            keep += [nbr, w]; nC[d] = nbr.ctypes.data; wC[d] = w.ctypes.data
            L, R, deg, nbr, w = graphs[(1, d)]
            LD[d], RD[d], degD = L, R, deg
            nbr = np.ascontiguousarray(nbr, dtype=np.uint32); w = np.ascontiguousarray(w, dtype=np.uint64)
            keep += [nbr, w]; nD[d] = nbr.ctypes.data; wD[d] = w.ctypes.data
        self._ck(self.lib.hb_expander_set(self.h, ctypes.c_longlong(n), levels, degC, degD, LC, RC, nC, wC, LD, RD, nD, wD))
        return int(self.lib.hb_expander_codeword_len(self.h))

    def encode(self, src, n, ncols):
        """src: (n, ncols) row-major messages-as-columns -> (2n, ncols)."""
        s = _F(src)
        d = np.zeros((2 * n * ncols, 2), dtype=np.uint64)
        self._ck(self.lib.hb_encode_batch(self.h, _ptr(s), _ptr(d), ctypes.c_longlong(n), c_sz(ncols)))
        return d

    # ---- hashes ----
    def blake3(self, src):
        s = np.ascontiguousarray(src, dtype=np.uint8).reshape(-1, 64)
        d = np.zeros((len(s), 32), dtype=np.uint8)
        self._ck(self.lib.hb_blake3_64(self.h, _ptr(s), _ptr(d), c_sz(len(s))))
        return d

    def create_tree(self, leaves):
        lv = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
        n = len(lv)
        out = np.zeros((2 * n - 1, 32), dtype=np.uint8)
        out[:n] = lv
        self._ck(self.lib.hb_merkle_tree(self.h, _ptr(out), c_sz(n)))
        return out

    def mt_commit_blake(self, leafs):
        x = _F(leafs)
        n = len(x) // 4
        out = np.zeros((2 * n - 1, 32), dtype=np.uint8)
        self._ck(self.lib.hb_mt_commit(self.h, _ptr(x), c_sz(len(x)), _ptr(out)))
        return out

    # ---- tensor code / commits ----
    def tensorcode(self, msg, trs, lin):
        m = _F(msg)
        out = np.zeros((4 * len(m), 2), dtype=np.uint64)
        self._ck(self.lib.hb_tensorcode(self.h, _ptr(m), c_sz(len(m)), int(trs), int(lin), _ptr(out)))
        return out

    def commit_standard(self, poly, K, trs, lin, want_tensor=False, levels_out=None, N=None):
        """poly: numpy (N,2) host array, or an int device pointer with N given."""
        if isinstance(poly, int):
            p = poly
        else:
            p = _F(poly); N = len(p)
        B = N // K
        levels = levels_out if levels_out is not None else np.zeros((2 * B - 1, 32), dtype=np.uint8)
        tensor = np.zeros((4 * N, 2), dtype=np.uint64) if want_tensor else None
        self._ck(self.lib.hb_commit_standard(self.h, _ptr(p), c_sz(N), int(K), int(trs), int(lin), _ptr(levels), _ptr(tensor)))
        return levels, tensor

    def tensor_gather(self, col, row, K):
        col = np.ascontiguousarray(col, dtype=np.uint32); row = np.ascontiguousarray(row, dtype=np.uint32)
        out = np.zeros((len(col) * K, 2), dtype=np.uint64)
        self._ck(self.lib.hb_tensor_gather(self.h, _ptr(col), _ptr(row), c_sz(len(col)), _ptr(out)))
        return out.reshape(len(col), K, 2)

    def aggregate(self, poly, K, beta, N=None):
        b = _F(beta)
        if poly is None:
            p = None
        else:
            p = _F(poly); N = len(p)
        out = np.zeros((N // K, 2), dtype=np.uint64)
        self._ck(self.lib.hb_aggregate(self.h, _ptr(p), c_sz(N), int(K), _ptr(b), _ptr(out)))
        return out

    def stream_pc_test(self, n):
        out = np.zeros((n, 2), dtype=np.uint64)
        self._ck(self.lib.hb_stream_pc_test(self.h, _ptr(out), c_sz(n)))
        return out

    def elastic_commit_async_levels(self, chunks, B, trs, lin):
        """Same commitment, the levels delivered level by level in the BACKGROUND (hb_elastic_finish_levels_async + hb_levels_wait)."""
        self._ck(self.lib.hb_elastic_begin(self.h, c_sz(B), int(trs), int(lin)))
        for c in chunks:
            self._ck(self.lib.hb_elastic_push(self.h, _ptr(_F(c))))
        nlev = int(np.log2(4 * B)) + 1
        lv = [np.empty((4 * B >> l, 32), dtype=np.uint8) for l in range(nlev)]
        ptrs = (c_vp * nlev)(*[x.ctypes.data_as(c_vp) for x in lv])
        self._ck(self.lib.hb_elastic_finish_levels_async(self.h, ptrs, nlev))
        self._ck(self.lib.hb_levels_wait(self.h))
        return np.concatenate(lv)

    def elastic_commit(self, chunks, B, trs, lin, reuse_pinned=False):
        """chunks: iterable of (B,2) arrays (the stream), pushed in order.  reuse_pinned: every chunk is first copied into ONE pinned host
        buffer that is overwritten right after the push returns — what a streaming producer does (hb_elastic_push returns only once the
        buffer has been read)."""
        self._ck(self.lib.hb_elastic_begin(self.h, c_sz(B), int(trs), int(lin)))
        pin = self.pinned((B, 2), np.uint64) if reuse_pinned else None
        for c in chunks:
            c = _F(c)
            assert len(c) == B
            if pin is not None:
                pin[:] = c
                self._ck(self.lib.hb_elastic_push(self.h, _ptr(pin)))
                pin[:] = 0xdeadbeef                                  # the producer refills its buffer at once
            else:
                self._ck(self.lib.hb_elastic_push(self.h, _ptr(c)))
        levels = np.zeros((8 * B - 1, 32), dtype=np.uint8)
        self._ck(self.lib.hb_elastic_finish(self.h, _ptr(levels)))
        return levels

    # ---- eq / MLE ----
    def precompute_beta(self, r):
        r = _F(r)
        out = np.zeros((1 << len(r), 2), dtype=np.uint64)
        self._ck(self.lib.hb_precompute_beta(self.h, _ptr(r), len(r), _ptr(out)))
        return out

    def evaluate_vector(self, v, r):
        v, r, out = _F(v), _F(r), np.zeros((1, 2), dtype=np.uint64)
        self._ck(self.lib.hb_evaluate_vector(self.h, _ptr(v), c_sz(len(v)), _ptr(r), _ptr(out)))
        return out

    # ---- sumchecks ----
    def sumcheck2(self, v1, v2, prev_r):
        v1, v2, r = _F(v1), _F(v2), _F(prev_r)
        rounds = int(np.log2(len(v1)))
        out = np.zeros((4 * rounds + 3, 2), dtype=np.uint64)
        ps = ctypes.c_double(0)
        self._ck(self.lib.hb_sumcheck2(self.h, _ptr(v1), _ptr(v2), c_sz(len(v1)), _ptr(r), _ptr(out), ctypes.byref(ps)))
        return out, ps.value

    def sumcheck3(self, v1, v2, v3, prev_r):
        v1, v2, v3, r = _F(v1), _F(v2), _F(v3), _F(prev_r)
        rounds = int(np.log2(len(v1)))
        out = np.zeros((5 * rounds + 4, 2), dtype=np.uint64)
        ps = ctypes.c_double(0)
        self._ck(self.lib.hb_sumcheck3(self.h, _ptr(v1), _ptr(v2), _ptr(v3), c_sz(len(v1)), _ptr(r), _ptr(out), ctypes.byref(ps)))
        return out, ps.value

    def batch_sumcheck3(self, t1, t2, t3, sizes, a):
        t1, t2, t3, a = _F(t1), _F(t2), _F(t3), _F(a)
        sz = (c_sz * len(sizes))(*sizes)
        rounds = int(np.log2(max(sizes)))
        out = np.zeros((5 * rounds + 3 * len(sizes), 2), dtype=np.uint64)
        ps = ctypes.c_double(0)
        self._ck(self.lib.hb_batch_sumcheck3(self.h, _ptr(t1), _ptr(t2), _ptr(t3), sz, len(sizes), _ptr(a), _ptr(out), ctypes.byref(ps)))
        return out, ps.value

    def mul_tree(self, inp, vectors, prev_r, x_rand=None):
        x, r = _F(inp), _F(prev_r)
        n = len(x) // vectors
        xr = _F(x_rand) if x_rand is not None else np.zeros((1, 2), dtype=np.uint64)
        out = np.zeros((16 + vectors + 8 * int(np.log2(len(x)) + 2) ** 2, 2), dtype=np.uint64)
        written, nfr, ps = c_sz(0), ctypes.c_int(0), ctypes.c_double(0)
        self._ck(self.lib.hb_mul_tree(self.h, _ptr(x), int(vectors), c_sz(n), _ptr(r), _ptr(xr), _ptr(out),
                                      ctypes.byref(written), ctypes.byref(nfr), ctypes.byref(ps)))
        return out[: written.value].copy(), nfr.value, ps.value

    # ---- streaming folding sumcheck (S4) / shallow streaming product tree (S6) ----
    def stream_layer(self, xy, B, layer_id, r, old_claim, rnd4):
        xy, r, oc, rnd = _F(xy), _F(r), _F(old_claim), _F(rnd4)
        nc, nr = np.zeros((1, 2), dtype=np.uint64), np.zeros((64, 2), dtype=np.uint64)
        n, ps = ctypes.c_int(0), ctypes.c_double(0)
        self._ck(self.lib.hb_stream_sumcheck_layer(self.h, _ptr(xy), c_sz(len(xy)), c_sz(B), int(layer_id), _ptr(r), len(r), _ptr(oc), _ptr(rnd),
                                                   _ptr(nc), _ptr(nr), ctypes.byref(n), ctypes.byref(ps)))
        return nc, nr[: n.value].copy(), ps.value

    def mul_tree_stream(self, xy, vectors, B, distance, naive, prev_r, x_rand, rnd):
        xy, pr, xr, rnd = _F(xy), _F(prev_r), _F(x_rand), _F(rnd)
        out = np.zeros((vectors, 2), dtype=np.uint64)
        layers, ps = ctypes.c_int(0), ctypes.c_double(0)
        self._ck(self.lib.hb_mul_tree_stream(self.h, _ptr(xy), c_sz(len(xy)), int(vectors), c_sz(B), int(distance), int(naive), _ptr(pr), _ptr(xr),
                                             _ptr(rnd), _ptr(out), ctypes.byref(layers), ctypes.byref(ps)))
        return out, ps.value, layers.value

    # ---- sharded commit building blocks (hobbit_b200/dist.py) ----
    def commit_encode_chunks(self, poly, nchunks, B, trs, lin, inner_out=None, leaf_parts=1, first_chunk=0, total_chunks=0):
        """poly / inner_out: numpy arrays or int device pointers.  Returns inner digests (nchunks*B, 32) when inner_out is None.
        leaf_parts > 1 writes the exchange layout [part][chunk][leaf in part] (see include/hobbit_b200.h)."""
        p = poly if isinstance(poly, int) else _F(poly)
        ret = None
        if inner_out is None:
            ret = np.zeros((nchunks * B, 32), dtype=np.uint8)
            inner_out = ret
        self._ck(self.lib.hb_commit_encode_chunks(self.h, _ptr(p), c_sz(nchunks), c_sz(B), int(trs), int(lin), _ptr(inner_out),
                                                  c_sz(leaf_parts), c_sz(first_chunk), c_sz(total_chunks)))
        return ret

    def elastic_encode_groups(self, chunks, ngroups, B, trs, lin, inner_out, leaf_parts=1, first_group=0, total_groups=0):
        """chunks: numpy F array or int device pointer (ngroups*4*B coefficients); inner_out: int device pointer or numpy uint8."""
        c = chunks if isinstance(chunks, int) else _F(chunks)
        self._ck(self.lib.hb_elastic_encode_groups(self.h, _ptr(c), c_sz(ngroups), c_sz(B), int(trs), int(lin), _ptr(inner_out),
                                                   c_sz(leaf_parts), c_sz(first_group), c_sz(total_groups)))

    def md_chain(self, inner, nchunks, nleaves, leaves):
        """leaves: numpy (nleaves, 32) uint8 or int device pointer; updated in place."""
        i = inner if isinstance(inner, int) else np.ascontiguousarray(inner, dtype=np.uint8)
        self._ck(self.lib.hb_md_chain(self.h, _ptr(i), c_sz(nchunks), c_sz(nleaves), _ptr(leaves)))
        return leaves

    def merkle_tree_inplace(self, levels, nleaves):
        """levels: int device pointer or numpy (2*nleaves-1, 32) with the leaves already at the front."""
        self._ck(self.lib.hb_merkle_tree(self.h, _ptr(levels), c_sz(nleaves)))
        return levels

    def gate_consistency(self, L, R, O, add, r):
        L, R, O, add, r = _F(L), _F(R), _F(O), _F(add), _F(r)
        rounds = int(np.log2(len(L)))
        out = np.zeros((6 * rounds + 6, 2), dtype=np.uint64)
        self._ck(self.lib.hb_gate_consistency_standard(self.h, _ptr(L), _ptr(R), _ptr(O), _ptr(add), c_sz(len(L)), _ptr(r), _ptr(out)))
        return out

    def gate_consistency_stream(self, L, R, O, S, B, r, rnd10):
        L, R, O, S, r, rnd = _F(L), _F(R), _F(O), _F(S), _F(r), _F(rnd10)
        cs = len(L); nch = cs // B
        lgB, lgn = int(np.log2(B)), int(np.log2(nch))
        out = np.zeros((nch + 6 * lgB + 6 + 6 * nch + 4 * lgn + 3, 2), dtype=np.uint64)
        ps = ctypes.c_double(0)
        self._ck(self.lib.hb_gate_consistency_stream(self.h, _ptr(L), _ptr(R), _ptr(O), _ptr(S), c_sz(cs), c_sz(B), _ptr(r), _ptr(rnd),
                                                     _ptr(out), ctypes.byref(ps)))
        return out, ps.value

    def elastic_open_front(self, chunks, betas, B, trs, lin, col, row):
        """O2 front: returns (agg (B,2), reply (queries, nchunks, 2))."""
        col = np.ascontiguousarray(col, dtype=np.uint32); row = np.ascontiguousarray(row, dtype=np.uint32)
        chunks = list(chunks); betas = _F(betas)
        self._ck(self.lib.hb_elastic_open_begin(self.h, c_sz(B), int(trs), int(lin), _ptr(col), _ptr(row), c_sz(len(col)), c_sz(len(chunks))))
        for i, c in enumerate(chunks):
            c = _F(c)
            b = np.ascontiguousarray(betas[i:i + 1])
            self._ck(self.lib.hb_elastic_open_push(self.h, _ptr(c), _ptr(b)))
        agg = np.zeros((B, 2), dtype=np.uint64)
        reply = np.zeros((len(col) * len(chunks), 2), dtype=np.uint64)
        self._ck(self.lib.hb_elastic_open_finish(self.h, _ptr(agg), _ptr(reply)))
        return agg, reply.reshape(len(col), len(chunks), 2)

    # ---- multi-GPU (include/hobbit_b200.h "multi-GPU"): one process per GPU ----
    def dist_local_info(self, data_bytes):
        blob = np.zeros(256, dtype=np.uint8)
        self._ck(self.lib.hb_dist_local_info(self.h, c_sz(int(data_bytes)), _ptr(blob)))
        return blob

    def dist_connect(self, rank, world, blobs):
        blobs = np.ascontiguousarray(np.asarray(blobs, dtype=np.uint8).reshape(world, 256))
        self._ck(self.lib.hb_dist_connect(self.h, int(rank), int(world), _ptr(blobs)))

    def dist_init_torch(self, data_bytes, group=None):
        """Bootstrap through an initialised torch.distributed process group (any backend): all-gathers the 256-byte window blobs."""
        import torch
        import torch.distributed as dist
        blob = self.dist_local_info(data_bytes)
        rank, world = dist.get_rank(group), dist.get_world_size(group)
        dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
        mine = torch.from_numpy(blob).to(dev)
        allb = [torch.empty_like(mine) for _ in range(world)]
        dist.all_gather(allb, mine, group=group)
        self.dist_connect(rank, world, np.stack([b.cpu().numpy() for b in allb]))
        dist.barrier(group)

    def dist_disconnect(self):
        self._ck(self.lib.hb_dist_disconnect(self.h))

    def dist_barrier(self):
        self._ck(self.lib.hb_dist_barrier(self.h))

    def dist_shard(self, on=True):
        self._ck(self.lib.hb_dist_shard(self.h, 1 if on else 0))

    def dist_stats(self):
        out = np.zeros(3, dtype=np.uint64)
        self.lib.hb_dist_stats(self.h, _ptr(out))
        return {"reductions": int(out[0]), "gathers": int(out[1]), "barriers": int(out[2])}

    def dist_allreduce(self, vec):
        v = _F(vec).copy()
        self._ck(self.lib.hb_dist_allreduce(self.h, _ptr(v), c_sz(len(v))))
        return v

    def dist_commit_standard(self, poly_local, K_total, B, trs, lin, levels_out=None):
        """poly_local: this rank's K_total/world consecutive chunks (numpy F array, DevF or raw device pointer)."""
        lv = np.zeros((2 * B - 1, 32), dtype=np.uint8) if levels_out is None else levels_out
        src = poly_local if isinstance(poly_local, (int, DevF)) else _F(poly_local)
        self._ck(self.lib.hb_dist_commit_standard(self.h, _ptr(src), c_sz(K_total), c_sz(B), int(trs), int(lin), _ptr(lv)))
        return lv

    def dist_elastic_commit(self, chunks_local, groups_total, B, trs, lin, levels_out=None):
        lv = np.zeros((8 * B - 1, 32), dtype=np.uint8) if levels_out is None else levels_out
        src = chunks_local if isinstance(chunks_local, (int, DevF)) else _F(chunks_local)
        self._ck(self.lib.hb_dist_elastic_commit(self.h, _ptr(src), c_sz(groups_total), c_sz(B), int(trs), int(lin), _ptr(lv)))
        return lv

    def sc3_round(self, ins, outs, L, rand):
        """ins/outs: 3 int device pointers each; returns the 4 cubic coefficients (4,2) of this slice."""
        r, co = _F(rand), np.zeros((4, 2), dtype=np.uint64)
        self._ck(self.lib.hb_sc3_round(self.h, c_vp(ins[0]), c_vp(ins[1]), c_vp(ins[2]), c_vp(outs[0]), c_vp(outs[1]), c_vp(outs[2]), c_sz(L),
                                       _ptr(r), _ptr(co)))
        return co
