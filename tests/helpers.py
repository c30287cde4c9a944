"""ctypes loaders for the two CHECKERS (test infrastructure, never the product):

* ``Checker("orc")`` -> oracle/libhobbit_oracle.so  (plain-C restatement, oracle/hobbit_oracle.c)
* ``Checker("ref")`` -> oracle/_ref/libhobbit_ref.so (the unmodified reference compiled in place
  by oracle/Makefile; exists wherever it was prebuilt — it travels to the GPU box as a binary).

Both export the same flat functions (prefix ``orc_`` / ``ref_``), so every parity test can be run
against either.  F arrays are numpy uint64 of shape (n, 2) = (real, img), the reference's 16-byte POD.
"""
import ctypes
import os
import subprocess

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P61 = (1 << 61) - 1
_libc = ctypes.CDLL(None)


def srand(seed=1):
    """glibc rand()/random() share one state; seed 1 == a fresh process."""
    _libc.srand(ctypes.c_uint(seed))


def F(arr):
    a = np.ascontiguousarray(np.asarray(arr, dtype=np.uint64))
    return a.reshape(-1, 2)


def fzeros(n):
    return np.zeros((n, 2), dtype=np.uint64)


def rand_field(rng, n, full=True):
    """Canonical random F_{p^2} elements (full-width limbs unless full=False -> small reals)."""
    if full:
        return rng.integers(0, P61, size=(n, 2), dtype=np.uint64)
    out = np.zeros((n, 2), dtype=np.uint64)
    out[:, 0] = rng.integers(0, 1 << 32, size=n, dtype=np.uint64)
    return out


def _p(a):
    return a.ctypes.data_as(ctypes.c_void_p)


def ref_available():
    return os.path.exists(os.path.join(ROOT, "oracle", "_ref", "libhobbit_ref.so"))


def ensure_oracle_built():
    so = os.path.join(ROOT, "oracle", "libhobbit_oracle.so")
    src = os.path.join(ROOT, "oracle", "hobbit_oracle.c")
    if not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle"], stdout=subprocess.DEVNULL)
    return so


class Checker:
    def __init__(self, kind):
        self.kind = kind
        if kind == "orc":
            self.lib = ctypes.CDLL(ensure_oracle_built())
            self.pfx = "orc_"
        elif kind == "ref":
            self.lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "_ref", "libhobbit_ref.so"))
            self.pfx = "ref_"
            self.lib.ref_init()
        else:
            raise ValueError(kind)

    def fn(self, name, restype=None):
        f = getattr(self.lib, self.pfx + name)
        f.restype = restype
        return f

    # ---- field ----
    def binop(self, op, a, b):
        a, b = F(a), F(b)
        c = np.empty_like(a)
        self.fn("field_binop")(op, _p(a), _p(b), _p(c), ctypes.c_size_t(len(a)))
        return c

    def root_of_unity(self, n):
        c = fzeros(1)
        self.fn("root_of_unity")(n, _p(c))
        return c

    def mimc(self, x, k):
        x, k, c = F(x), F(k), fzeros(1)
        self.fn("mimc_hash")(_p(x), _p(k), _p(c))
        return c

    def fft(self, arr, logn):
        a = F(arr).copy()
        self.fn("fft")(_p(a), logn)
        return a

    def generate_randomness(self, n):
        c = fzeros(n)
        self.fn("generate_randomness")(n, _p(c))
        return c

    # ---- expander ----
    def expander_init_store(self, n):
        return self.fn("expander_init_store", ctypes.c_longlong)(ctypes.c_longlong(n))

    def expander_graphs(self, n):
        """[(which, dep, L, R, deg, nbr[L*deg] u32, w[L*deg] u64)] in generation order is not needed; index by (which,dep)."""
        out = {}
        for dep in range(self.fn("expander_levels", ctypes.c_int)(ctypes.c_longlong(n))):
            for which in (0, 1):
                R, deg = ctypes.c_longlong(), ctypes.c_int()
                L = self.fn("expander_dims", ctypes.c_longlong)(which, dep, ctypes.byref(R), ctypes.byref(deg))
                nbr = np.zeros(L * deg.value, dtype=np.uint32)
                w = np.zeros(L * deg.value, dtype=np.uint64)
                self.fn("expander_dump")(which, dep, _p(nbr), _p(w))
                out[(which, dep)] = (L, R.value, deg.value, nbr, w)
        return out

    def encode(self, src, n):
        s = F(src)
        d = fzeros(2 * n)
        cw = self.fn("encode_monolithic", ctypes.c_int)(_p(s), _p(d), ctypes.c_longlong(n))
        return d, cw

    def encode_reseed(self, src, n):
        """E3: encode() — graph re-drawn from fixed seeds per call; returns the n + L + R codeword entries."""
        s = F(src)
        d = fzeros(2 * n)
        cw = self.fn("encode_reseed", ctypes.c_int)(_p(s), _p(d), ctypes.c_longlong(n))
        return d[:cw].copy(), cw

    # ---- hashes ----
    def blake3(self, src64):
        s = np.ascontiguousarray(src64, dtype=np.uint8)
        d = np.zeros(32, dtype=np.uint8)
        self.fn("blake3_hash")(_p(s), _p(d))
        return d

    def md_leaf(self, xyzw, prev):
        x, p = F(xyzw), np.ascontiguousarray(prev, dtype=np.uint8)
        d = np.zeros(32, dtype=np.uint8)
        self.fn("md_leaf")(_p(x), _p(p), _p(d))
        return d

    def mt_commit_blake(self, leafs):
        x = F(leafs)
        n = len(x) // 4
        out = np.zeros((2 * n - 1, 32), dtype=np.uint8)
        self.fn("mt_commit_blake", ctypes.c_int)(_p(x), len(x), _p(out))
        return out

    def create_tree(self, leaves):
        lv = np.ascontiguousarray(leaves, dtype=np.uint8).reshape(-1, 32)
        n = len(lv)
        out = np.zeros((2 * n - 1, 32), dtype=np.uint8)
        self.fn("create_tree_blake", ctypes.c_int)(_p(lv), n, _p(out))
        return out

    # ---- tensor code / commits ----
    def tensorcode(self, msg, trs, lin):
        m = F(msg)
        out = fzeros(4 * len(m))
        self.fn("compute_tensorcode")(_p(m), ctypes.c_size_t(len(m)), trs, int(lin), _p(out))
        return out

    def commit_standard(self, poly, K, trs, lin, want_tensor=False):
        p = F(poly)
        N = len(p)
        B = N // K
        levels = np.zeros((2 * B - 1, 32), dtype=np.uint8)
        tensor = fzeros(4 * N) if want_tensor else None
        self.fn("commit_standard", ctypes.c_double)(_p(p), ctypes.c_size_t(N), K, trs, int(lin), _p(levels),
                                                    _p(tensor) if want_tensor else None)
        return levels, tensor

    def read_stream_pc_test(self, n):
        out = fzeros(n)
        self.fn("read_stream_pc_test")(_p(out), ctypes.c_size_t(n))
        return out

    def elastic_commit(self, N, B, trs, lin):
        levels = np.zeros((8 * B - 1, 32), dtype=np.uint8)
        self.fn("elastic_commit", ctypes.c_double)(ctypes.c_size_t(N), ctypes.c_size_t(B), trs, int(lin), _p(levels))
        return levels

    # ---- eq table / MLE ----
    def precompute_beta(self, r):
        r = F(r)
        out = fzeros(1 << len(r))
        self.fn("precompute_beta")(_p(r), len(r), _p(out))
        return out

    def evaluate_vector(self, v, r):
        v, r, out = F(v), F(r), fzeros(1)
        self.fn("evaluate_vector")(_p(v), ctypes.c_size_t(len(v)), _p(r), len(r), _p(out))
        return out

    # ---- sumchecks: return (flat proof array, ps) ----
    def sumcheck2(self, v1, v2, prev_r):
        v1, v2, r = F(v1), F(v2), F(prev_r)
        rounds = int(np.log2(len(v1)))
        out = fzeros(4 * rounds + 3)
        ps = self.fn("sumcheck2", ctypes.c_double)(_p(v1), _p(v2), ctypes.c_size_t(len(v1)), _p(r), _p(out))
        return out, ps

    def sumcheck3(self, v1, v2, v3, prev_r):
        v1, v2, v3, r = F(v1), F(v2), F(v3), F(prev_r)
        rounds = int(np.log2(len(v1)))
        out = fzeros(5 * rounds + 4)
        ps = self.fn("sumcheck3", ctypes.c_double)(_p(v1), _p(v2), _p(v3), ctypes.c_size_t(len(v1)), _p(r), _p(out))
        return out, ps

    def batch_sumcheck3(self, t1, t2, t3, sizes, a):
        t1, t2, t3, a = F(t1), F(t2), F(t3), F(a)
        sz = (ctypes.c_size_t * len(sizes))(*sizes)
        rounds = int(np.log2(max(sizes)))
        out = fzeros(5 * rounds + 3 * len(sizes))
        ps = self.fn("batch_sumcheck3", ctypes.c_double)(_p(t1), _p(t2), _p(t3), sz, len(sizes), _p(a), _p(out))
        return out, ps

    def mul_tree(self, inp, vectors, prev_r):
        x, r = F(inp), F(prev_r)
        n = len(x) // vectors
        out = fzeros(16 + vectors + 8 * int(np.log2(len(x)) + 2) ** 2)
        nfr, ps = ctypes.c_int(), ctypes.c_double()
        k = self.fn("mul_tree", ctypes.c_size_t)(_p(x), vectors, ctypes.c_size_t(n), _p(r), _p(out),
                                                 ctypes.byref(nfr), ctypes.byref(ps))
        return out[:k].copy(), nfr.value, ps.value


def synthetic_stream(total):
    """read_stream's default branch (witness_stream.cpp:2348-2352): v[i] = F(i%1024 + 1), restarted on every read.  In the
    logical two-half form [X | Y] both halves are the same periodic sequence as long as every read is a multiple of 2048
    elements, i.e. BUFFER_SPACE >= 512 (smaller buffers make the reference's synthetic stream depend on the read size)."""
    out = np.zeros((total, 2), dtype=np.uint64)
    out[:, 0] = (np.arange(total, dtype=np.uint64) % 1024) + 1
    return out


def _stream_batch(self, xy, B, layer, distance, batches, r_rows, old_claims):
    """batched S4; r_rows: list of (len_j, 2) arrays.  Returns (new_claims (batches,2), [new_r rows], ps)."""
    xy = F(xy)
    rs = max(len(x) for x in r_rows) + 4
    r = np.zeros((batches, rs, 2), dtype=np.uint64)
    rlen = (ctypes.c_int * batches)(*[len(x) for x in r_rows])
    for j, x in enumerate(r_rows):
        r[j, :len(x)] = x
    oc = F(old_claims)
    nc, nr = fzeros(batches), np.zeros((batches, rs, 2), dtype=np.uint64)
    ps = ctypes.c_double(0)
    if self.kind == "orc":
        n0 = self.fn("stream_sumcheck_batch", ctypes.c_int)(_p(xy), ctypes.c_size_t(len(xy)), ctypes.c_size_t(B), layer, distance, batches, _p(r), rs, rlen,
                                                           _p(oc), _p(nc), _p(nr), ctypes.byref(ps))
    else:
        n0 = self.fn("stream_sumcheck_batch", ctypes.c_int)(ctypes.c_size_t(len(xy)), ctypes.c_size_t(B), layer, distance, batches, _p(r), rs, rlen,
                                                           _p(oc), _p(nc), _p(nr), ctypes.byref(ps))
    rows = [nr[j, :n0 - j * distance].copy() for j in range(batches)]
    return nc, rows, ps.value


Checker.stream_batch = _stream_batch


def _stream_layer(self, xy, B, layer_id, r, old_claim):
    """S4.  Checker('ref') ignores xy (it reads the synthetic stream itself)."""
    xy, r, oc = F(xy), F(r), F(old_claim)
    total = len(xy)
    nc, nr = fzeros(1), fzeros(64)
    ps = ctypes.c_double()
    if self.kind == "ref":
        n = self.fn("stream_sumcheck_layer", ctypes.c_int)(ctypes.c_size_t(total), ctypes.c_size_t(B), layer_id, _p(r), len(r), _p(oc),
                                                           _p(nc), _p(nr), ctypes.byref(ps))
    else:
        n = self.fn("stream_sumcheck_layer", ctypes.c_int)(_p(xy), ctypes.c_size_t(total), ctypes.c_size_t(B), layer_id, _p(r), len(r), _p(oc),
                                                           _p(nc), _p(nr), ctypes.byref(ps))
    return nc, nr[:n].copy(), ps.value


def _mul_tree_stream(self, xy, vectors, B, distance, naive, prev_r):
    xy, pr = F(xy), F(prev_r)
    out = fzeros(vectors)
    if self.kind == "ref":
        ps = self.fn("mul_tree_stream", ctypes.c_double)(ctypes.c_size_t(len(xy)), vectors, ctypes.c_size_t(B), distance, naive, _p(pr), _p(out))
    else:
        ps = self.fn("mul_tree_stream", ctypes.c_double)(_p(xy), ctypes.c_size_t(len(xy)), vectors, ctypes.c_size_t(B), distance, naive, _p(pr), _p(out))
    return out, ps


Checker.stream_layer = _stream_layer
Checker.mul_tree_stream = _mul_tree_stream


def _gate_consistency(self, L, R, O, add, r):
    """Returns (final add, L, R, O) [4 F] for 'ref'; the full transcript (6*rounds + 6 F) for 'orc'."""
    L, R, O, add, r = F(L), F(R), F(O), F(add), F(r)
    n = len(L)
    rounds = int(np.log2(n))
    if self.kind == "ref":
        out = fzeros(4)
        self.fn("gate_consistency_standard")(_p(L), _p(R), _p(O), _p(add), ctypes.c_size_t(n), _p(r), _p(out))
        return out
    out = fzeros(6 * rounds + 6)
    self.fn("gate_consistency_standard")(_p(L), _p(R), _p(O), _p(add), ctypes.c_size_t(n), _p(r), _p(out))
    return out


Checker.gate_consistency = _gate_consistency


def _gate_stream(self, L, R, O, S, B, r):
    assert self.kind == "orc", "the reference's prove_gate_consistency needs its circuit evaluator thread"
    L, R, O, S, r = F(L), F(R), F(O), F(S), F(r)
    cs = len(L); nch = cs // B
    lgB, lgn = int(np.log2(B)), int(np.log2(nch))
    out = fzeros(nch + 6 * lgB + 6 + 6 * nch + 4 * lgn + 3)
    ps = self.fn("gate_consistency_stream", ctypes.c_double)(_p(L), _p(R), _p(O), _p(S), ctypes.c_size_t(cs), ctypes.c_size_t(B), _p(r), _p(out))
    return out, ps


Checker.gate_stream = _gate_stream


def consistent_trace(rng, orc, cs):
    """A gate transcript that satisfies O = S ? L+R : L*R, i.e. what read_trace emits for a correctly evaluated circuit."""
    L, R = rand_field(rng, cs), rand_field(rng, cs)
    S = np.zeros((cs, 2), dtype=np.uint64)
    S[:, 0] = rng.integers(0, 2, cs)
    O = np.where(S[:, :1] == 1, orc.binop(0, L, R), orc.binop(2, L, R))
    return L, R, O, S


def _elastic_open_front(self, stream, B, trs, lin, beta, col, row):
    """O2 front.  'ref': RS columns only, reads its own synthetic stream (stream is ignored)."""
    stream, beta = F(stream), F(beta)
    col = np.ascontiguousarray(col, dtype=np.uint32); row = np.ascontiguousarray(row, dtype=np.uint32)
    nch = len(stream) // B
    agg, reply = fzeros(B), fzeros(len(col) * nch)
    if self.kind == "ref":
        assert not lin
        self.fn("elastic_open_front_rs")(ctypes.c_size_t(len(stream)), ctypes.c_size_t(B), trs, _p(beta), _p(col), _p(row), ctypes.c_size_t(len(col)),
                                         _p(agg), _p(reply))
    else:
        self.fn("elastic_open_front")(_p(stream), ctypes.c_size_t(nch), ctypes.c_size_t(B), trs, int(lin), _p(beta), _p(col), _p(row),
                                      ctypes.c_size_t(len(col)), _p(agg), _p(reply))
    return agg, reply.reshape(len(col), nch, 2)


Checker.elastic_open_front = _elastic_open_front


def synthetic_chunks(nch, B):
    """read_stream's default stream, chunk by chunk: every read restarts at i = 0 (witness_stream.cpp:2348-2352)."""
    c = np.zeros((B, 2), dtype=np.uint64)
    c[:, 0] = (np.arange(B, dtype=np.uint64) % 1024) + 1
    return np.concatenate([c] * nch)


# ---- the 8f.1 building blocks through the raw C ABI: the same ctypes calls against the CUDA library and against the CPU emulation
# (oracle/libhb_emul.so, test infrastructure), so each kernel is compared with its plain restatement -----------------------------
class RawABI:
    def __init__(self, kind):
        if kind == "gpu":
            path = os.path.join(ROOT, "hobbit_b200", "libhobbit_b200.so")
        elif kind == "emul":
            path = os.path.join(ROOT, "oracle", "libhb_emul.so")
            src = [os.path.join(ROOT, "oracle", f) for f in ("hb_emul.cpp", "hobbit_oracle.c")]
            if not os.path.exists(path) or any(os.path.getmtime(path) < os.path.getmtime(s) for s in src):
                subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle"), "liboracle"], stdout=subprocess.DEVNULL)
        else:
            raise ValueError(kind)
        self.lib = ctypes.CDLL(path)
        self.lib.hb_last_error.restype = ctypes.c_char_p
        self.ctx = ctypes.c_void_p()
        if self.lib.hb_ctx_create(ctypes.byref(self.ctx), 0):
            raise RuntimeError("hb_ctx_create failed")

    def call(self, name, *args):
        conv = []
        for a in args:
            if isinstance(a, np.ndarray):
                conv.append(_p(a))
            elif isinstance(a, int):
                conv.append(ctypes.c_size_t(a))
            else:
                conv.append(a)
        rc = getattr(self.lib, name)(self.ctx, *conv)
        if rc:
            raise RuntimeError("%s: %s" % (name, self.lib.hb_last_error(self.ctx).decode()))

    def rs_encode_rows(self, src, in_len, rows, logn):
        out = fzeros(rows << logn)
        self.call("hb_rs_encode_rows", F(src), in_len, rows, out, ctypes.c_int(logn))
        return out

    def matvec_cols(self, M, rows, cols, w):
        out = fzeros(cols)
        self.call("hb_matvec_cols", F(M), rows, cols, cols, F(w), out)
        return out

    def matvec_rows(self, M, rows, cols, s):
        out = fzeros(rows)
        self.call("hb_matvec_rows", F(M), rows, cols, cols, F(s), out)
        return out

    def axpy(self, y, x, a):
        y = F(y).copy()
        self.call("hb_axpy", y, F(x), F(a), len(y))
        return y

    def scatter(self, n, idx, val):
        out = fzeros(n)
        idx = np.ascontiguousarray(idx, dtype=np.uint64)
        self.call("hb_scatter", out, n, idx, F(val), len(idx))
        return out

    def gather_cols(self, M, rows, cols, col):
        col = np.ascontiguousarray(col, dtype=np.uint64)
        out = fzeros(len(col) * rows)
        self.call("hb_gather_cols", F(M), rows, cols, cols, col, len(col), out)
        return out

    def phi_g_init(self, r):
        r = F(r)
        out = fzeros(1 << len(r))
        self.call("hb_phi_g_init", r, ctypes.c_int(len(r)), out)
        return out

    def shockwave_leaves(self, enc, k, cols):
        out = np.zeros((cols, 32), dtype=np.uint8)
        self.call("hb_shockwave_leaves", F(enc), ctypes.c_int(k), cols, out)
        return out

    def change_form(self, poly):
        p = F(poly).copy()
        self.call("hb_change_form", p, ctypes.c_int(int(np.log2(len(p)))))
        return p

    def regroup(self, v, k):
        v = F(v)
        out = fzeros(len(v))
        self.call("hb_regroup", v, len(v), ctypes.c_int(k), out)
        return out

    def whir_poly(self, poly, beta, L):
        out = fzeros(3)
        self.call("hb_whir_poly", F(poly), F(beta), L, out)
        return out

    def whir_fold(self, poly, beta, L, a):
        p, b = F(poly).copy(), F(beta).copy()
        self.call("hb_whir_fold", p, b, L, F(a))
        return p, b

    def whir_zeta(self, poly, beta, zetas, pows):
        poly, b = F(poly), F(beta).copy()
        v = int(np.log2(len(poly)))
        pows = F(pows)
        y = fzeros(len(pows))
        self.call("hb_whir_zeta", poly, b, ctypes.c_int(v), F(zetas), ctypes.c_int(len(pows)), pows, y)
        return b, y

    def select_cols(self, M, rows, cols, col):
        col = np.ascontiguousarray(col, dtype=np.uint64)
        out = fzeros(len(col) * rows)
        self.call("hb_select_cols", F(M), rows, cols, cols, col, len(col), out)
        return out

    def transpose(self, M, rows, cols):
        out = fzeros(rows * cols)
        self.call("hb_transpose", F(M), rows, cols, out)
        return out

    def any_nonzero(self, v):
        v = F(v)
        flag = ctypes.c_int(7)
        self.call("hb_any_nonzero", v, len(v), ctypes.byref(flag))
        return flag.value

    # ---- W1/W2: streams from a trace (records: numpy structured array with the 80-byte tr_tuple layout) ----
    def trace_load(self, records, split=3):
        self.call("hb_trace_begin", 0)
        done = ctypes.c_int(0)
        n = len(records)
        step = max(1, n // split)
        for off in range(0, n, step):
            part = np.ascontiguousarray(records[off:off + step])
            self.call("hb_trace_push", part.ctypes.data_as(ctypes.c_void_p), len(part), ctypes.byref(done))
        cnt = (ctypes.c_size_t * 3)()
        self.call("hb_trace_finish", ctypes.byref(cnt, 0), ctypes.byref(cnt, 8), ctypes.byref(cnt, 16))
        return tuple(cnt), done.value

    def trace_streams(self, cs, a_w, b_w, has_lookups=0):
        w, xy = fzeros(4 * cs), fzeros(8 * cs)
        L, R, O, S = fzeros(cs), fzeros(cs), fzeros(cs), fzeros(cs)
        self.call("hb_trace_witness", cs, w)
        self.call("hb_trace_transcript", cs, ctypes.c_int(has_lookups), L, R, O, S)
        self.call("hb_trace_wiring", cs, F(a_w), F(b_w), xy)
        return w, L, R, O, S, xy

    def trace_lookup_streams(self, cs, lookup_rand4):
        lb, lw = fzeros(2 * cs), fzeros(2 * cs)
        self.call("hb_trace_lookup_basic", cs, F(lookup_rand4), lb)
        self.call("hb_trace_lookup_witness", cs, F(lookup_rand4), lw)
        return lb, lw

    def binop(self, op, a, b):
        a, b = F(a), F(b)
        c = np.empty_like(a)
        self.call("hb_field_binop", ctypes.c_int(op), a, b, c, len(a))
        return c

    def gate_consistency_lookups(self, L, R, O, S, B, r, lookup_rand2, rnd13):
        L, R, O, S = F(L), F(R), F(O), F(S)
        cs = len(L); nch = cs // B
        lgB, lgn = int(np.log2(B)), int(np.log2(nch))
        out = fzeros(nch + 6 * lgB + 9 + 8 * nch + 4 * lgn + 3)
        ps = ctypes.c_double(0)
        self.call("hb_gate_consistency_lookups_stream", L, R, O, S, cs, B, F(r), F(lookup_rand2), F(rnd13), out, ctypes.byref(ps))
        return out, ps.value


TR_TUPLE = np.dtype([("value_o", np.uint64, 2), ("value_l", np.uint64, 2), ("value_r", np.uint64, 2), ("idx_o", np.int32), ("idx_l", np.int32),
                     ("idx_r", np.int32), ("access_o", np.int32), ("access_l", np.int32), ("access_r", np.int32), ("type", np.uint8), ("pad", np.uint8, 7)])
assert TR_TUPLE.itemsize == 80


def synthetic_trace(rng, n, lookups=False):
    """n records of a made-up trace (types 0/1/2, optionally 3..5) followed by the end marker and some garbage that must be ignored."""
    t = np.zeros(n + 5, dtype=TR_TUPLE)
    for f in ("value_o", "value_l", "value_r"):
        t[f] = rng.integers(0, P61, size=(n + 5, 2), dtype=np.uint64)
    for f in ("idx_o", "idx_l", "idx_r"):
        t[f] = rng.integers(0, 1 << 20, size=n + 5)
    for f in ("access_o", "access_l", "access_r"):
        t[f] = rng.integers(0, 50, size=n + 5)
    t["type"] = rng.choice([0, 1, 2, 3, 4, 5] if lookups else [0, 1, 2], size=n + 5)
    t["type"][n] = 255
    if lookups:                      # table entries are small integers (value_l, or value_l + 256 value_r): many repeats -> access counters
        lk = t["type"] >= 3
        t["value_l"][lk] = 0
        t["value_r"][lk] = 0
        t["value_l"][lk, 0] = rng.integers(0, 40, size=int(lk.sum()))
        t["value_r"][lk, 0] = rng.integers(0, 3, size=int(lk.sum()))
    return t


def _raw_mul_tree_stream(self, xy, vectors, B, distance, naive, prev_r, x_rand, rnd):
    xy = F(xy)
    out = fzeros(vectors)
    layers, ps = ctypes.c_int(0), ctypes.c_double(0)
    self.call("hb_mul_tree_stream", xy, len(xy), ctypes.c_int(vectors), B, ctypes.c_int(distance), ctypes.c_int(naive), F(prev_r), F(x_rand), F(rnd), out,
              ctypes.byref(layers), ctypes.byref(ps))
    return out, ps.value, layers.value


RawABI.mul_tree_stream = _raw_mul_tree_stream


def _raw_trace_generate_mlp(self, layer_size):
    ls = (ctypes.c_int * len(layer_size))(*layer_size)
    n = ctypes.c_size_t(0)
    self.call("hb_trace_generate_mlp", ls, ctypes.c_int(len(layer_size)), ctypes.byref(n))
    cnt = (ctypes.c_size_t * 3)()
    self.call("hb_trace_finish", ctypes.byref(cnt, 0), ctypes.byref(cnt, 8), ctypes.byref(cnt, 16))
    return tuple(cnt)


RawABI.trace_generate_mlp = _raw_trace_generate_mlp


def _raw_trace_generate_aes(self, input_size):
    n = ctypes.c_size_t(0)
    self.call("hb_trace_generate_aes", ctypes.c_int(input_size), ctypes.byref(n))
    cnt = (ctypes.c_size_t * 3)()
    self.call("hb_trace_finish", ctypes.byref(cnt, 0), ctypes.byref(cnt, 8), ctypes.byref(cnt, 16))
    assert cnt[0] == n.value
    return tuple(cnt)


RawABI.trace_generate_aes = _raw_trace_generate_aes


def _raw_trace_generate_sql(self, input_size):
    n = ctypes.c_size_t(0)
    self.call("hb_trace_generate_sql", ctypes.c_int(input_size), ctypes.byref(n))
    cnt = (ctypes.c_size_t * 3)()
    self.call("hb_trace_finish", ctypes.byref(cnt, 0), ctypes.byref(cnt, 8), ctypes.byref(cnt, 16))
    assert cnt[0] == n.value
    return tuple(cnt)


RawABI.trace_generate_sql = _raw_trace_generate_sql


def _raw_trace_generate_pruned(self, n_inputs, rows0, rows1):
    """rows_l: list (one entry per neuron) of lists of column indices"""
    def csr(rows):
        rp = np.zeros(len(rows) + 1, dtype=np.int32)
        rp[1:] = np.cumsum([len(r) for r in rows])
        cols = np.array([c for r in rows for c in r] or [0], dtype=np.int32)
        return rp, cols
    rp0, c0 = csr(rows0)
    rp1, c1 = csr(rows1)
    n = ctypes.c_size_t(0)
    self.call("hb_trace_generate_pruned_mlp", ctypes.c_int(n_inputs), ctypes.c_int(len(rows0)), ctypes.c_int(len(rows1)), rp0, c0, rp1, c1, ctypes.byref(n))
    cnt = (ctypes.c_size_t * 3)()
    self.call("hb_trace_finish", ctypes.byref(cnt, 0), ctypes.byref(cnt, 8), ctypes.byref(cnt, 16))
    assert cnt[0] == n.value
    return tuple(cnt)


RawABI.trace_generate_pruned = _raw_trace_generate_pruned
