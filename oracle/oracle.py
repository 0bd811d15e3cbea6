"""ctypes front end of the CPU ORACLE (oracle/kmerlr_oracle.c).

TEST INFRASTRUCTURE ONLY: imported by tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs.  The product package (kmerlr_b200/) must never import it.
"""
import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libkmerlr_oracle.so")

TIE_GO118, TIE_INDEX = 0, 1
SUMMARY = {"": 0, "mean": 1, "product": 2, "min": 3, "max": 4}


def build(force=False):
    src = [os.path.join(_HERE, f) for f in ("kmerlr_oracle.c", "kmerlr_oracle.h")]
    if force or not os.path.exists(_SO) or any(os.path.getmtime(s) > os.path.getmtime(_SO) for s in src):
        subprocess.check_call(["make", "-s", "-C", _HERE, "libkmerlr_oracle.so"])
    return _SO


class Config(C.Structure):
    _fields_ = [(n, C.c_int32) for n in
                ("M", "N", "complement", "reverse", "revcomp", "binarize", "alphabet", "max_ambiguous")]


def make_config(M, N, complement=False, reverse=False, revcomp=False, binarize=False, alphabet="nucleotide",
                max_ambiguous=-1):
    return Config(M, N, int(complement), int(reverse), int(revcomp), int(binarize),
                  {"nucleotide": 0, "gapped-nucleotide": 1}[alphabet], max_ambiguous)


class HookState(C.Structure):
    _fields_ = [("loss_old", C.c_double), ("loss_new", C.c_double)]


class Estimator(C.Structure):
    _fields_ = [("n_active", C.c_int64), ("active_idx", C.POINTER(C.c_int64)), ("theta", C.POINTER(C.c_double)),
                ("state_cap", C.c_int64), ("hook", HookState), ("l1reg_over_n", C.c_double)]


class Model(C.Structure):
    _fields_ = [("cfg", Config), ("n_classes", C.c_int64), ("class_k", C.POINTER(C.c_int32)),
                ("class_code", C.POINTER(C.c_uint64)), ("n_features", C.c_int64),
                ("features", C.POINTER(C.c_int32)), ("n_members", C.c_int64), ("theta", C.POINTER(C.c_double)),
                ("summary", C.c_int32)]


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, i64, i32, dbl = C.c_void_p, C.c_int64, C.c_int32, C.c_double
        L.ko_extract.restype = vp
        L.ko_extract.argtypes = [C.POINTER(Config), vp, vp, i64, vp, vp, i64, vp, i64, C.c_int, C.c_int]
        L.ko_matrix_free.argtypes = [vp]
        L.ko_matrix_info.argtypes = [vp] + [C.POINTER(i64)] * 4
        L.ko_matrix_classes.argtypes = [vp, vp, vp]
        L.ko_matrix_rows.argtypes = [vp, vp, vp, vp]
        L.ko_matrix_from_csr.restype = vp
        L.ko_matrix_from_csr.argtypes = [i64, i64, vp, vp, vp]
        L.ko_class_name.restype = C.c_int
        L.ko_class_name.argtypes = [C.POINTER(Config), i32, C.c_uint64, C.c_char_p, C.c_int]
        L.ko_class_code.restype = C.c_uint64
        L.ko_class_code.argtypes = [C.POINTER(Config), vp, i32]
        L.ko_coeff_dim.restype = i64
        L.ko_coeff_dim.argtypes = [i64]
        L.ko_coeff_ind2sub.restype = i64
        L.ko_coeff_ind2sub.argtypes = [i64, i64, i64]
        L.ko_coeff_sub2ind.argtypes = [i64, i64, C.POINTER(i64), C.POINTER(i64)]
        L.ko_linear_pdf.argtypes = [vp, vp, C.c_int, vp]
        L.ko_log_pdf.argtypes = [vp, vp, C.c_int, vp]
        L.ko_gradient.argtypes = [vp, vp, vp, i64, vp, dbl, C.c_int, vp]
        L.ko_loss.restype = dbl
        L.ko_loss.argtypes = [vp, vp, vp, i64, vp, dbl, C.c_int]
        L.ko_class_weights.argtypes = [vp, i64, vp]
        L.ko_nlargest_abs.argtypes = [vp, vp, i64, C.c_int]
        L.ko_select.restype = C.c_int
        L.ko_select.argtypes = [vp, vp, vp, C.c_int, i64, dbl, vp, vp, i64, C.c_int, dbl, dbl, vp, i64,
                                C.POINTER(dbl), C.POINTER(i64), vp]
        L.ko_reduce.restype = vp
        L.ko_reduce.argtypes = [vp, vp, i64]
        L.ko_step_size.restype = dbl
        L.ko_step_size.argtypes = [vp, dbl, dbl]
        L.ko_proxgrad.restype = i64
        L.ko_proxgrad.argtypes = [vp, vp, vp, vp, dbl, dbl, dbl, dbl, dbl, i64, C.POINTER(HookState), C.POINTER(dbl)]
        L.ko_estimate_loop.restype = i64
        L.ko_estimate_loop.argtypes = [vp, vp, vp, C.c_int, i64, C.c_int, dbl, dbl, dbl, dbl, dbl, i64, i64,
                                       C.POINTER(Estimator), vp, vp, i64]
        L.ko_window_slots.restype = i64
        L.ko_window_slots.argtypes = [i64, i64, i64]
        L.ko_score_windows.argtypes = [C.POINTER(Model), C.c_int, vp, vp, i64, i64, i64, vp, C.c_int]
        _lib = L
    return _lib


def _p(a):
    return a.ctypes.data_as(C.c_void_p) if a is not None else None


def flatten(seqs):
    """list of bytes/str -> (uint8 concatenation, int64 offsets[n+1])"""
    bs = [s.encode() if isinstance(s, str) else bytes(s) for s in seqs]
    off = np.zeros(len(bs) + 1, dtype=np.int64)
    off[1:] = np.cumsum([len(b) for b in bs])
    buf = np.frombuffer(b"".join(bs), dtype=np.uint8).copy() if off[-1] else np.zeros(0, dtype=np.uint8)
    return buf, off


class Matrix:
    """KmerDataSet restated: rows (CSR without the bias column) + class list."""

    def __init__(self, handle, cfg=None):
        self.h = handle
        self.cfg = cfg
        n, m, nnz, nc = (C.c_int64() for _ in range(4))
        lib().ko_matrix_info(handle, n, m, nnz, nc)
        self.n, self.m, self.nnz, self.n_classes = n.value, m.value, nnz.value, nc.value

    def __del__(self):
        if getattr(self, "h", None):
            lib().ko_matrix_free(self.h)
            self.h = None

    def classes(self):
        k = np.zeros(self.n_classes, dtype=np.int32)
        code = np.zeros(self.n_classes, dtype=np.uint64)
        lib().ko_matrix_classes(self.h, _p(k), _p(code))
        return k, code

    def rows(self):
        rowptr = np.zeros(self.n + 1, dtype=np.int64)
        col = np.zeros(max(self.nnz, 1), dtype=np.int32)
        val = np.zeros(max(self.nnz, 1), dtype=np.float64)
        lib().ko_matrix_rows(self.h, _p(rowptr), _p(col), _p(val))
        return rowptr, col[:self.nnz], val[:self.nnz]

    def dense(self):
        rowptr, col, val = self.rows()
        d = np.zeros((self.n, self.m))
        for i in range(self.n):
            d[i, col[rowptr[i]:rowptr[i + 1]]] = val[rowptr[i]:rowptr[i + 1]]
        return d

    def class_names(self):
        k, code = self.classes()
        return [class_name(self.cfg, int(a), int(b)) for a, b in zip(k, code)]


def class_name(cfg, k, code):
    buf = C.create_string_buffer(256)
    lib().ko_class_name(C.byref(cfg), k, code, buf, 256)
    return buf.value.decode()


def extract(cfg, seqs, frozen=None, features=None, threads=0, faithful=False):
    buf, off = flatten(seqs) if not isinstance(seqs, tuple) else seqs
    fk = fc = None
    nf = 0
    if frozen is not None:
        fk = np.ascontiguousarray(frozen[0], dtype=np.int32)
        fc = np.ascontiguousarray(frozen[1], dtype=np.uint64)
        nf = len(fk)
    ft = None
    nft = 0
    if features is not None and len(features):
        ft = np.ascontiguousarray(features, dtype=np.int32).reshape(-1, 2)
        nft = ft.shape[0]
    h = lib().ko_extract(C.byref(cfg), _p(buf), _p(off), len(off) - 1, _p(fk), _p(fc), nf, _p(ft), nft,
                         threads, int(faithful))
    if not h:
        raise ValueError("ko_extract: bad configuration")
    return Matrix(h, cfg)


def from_csr(n, m, rowptr, col, val):
    rowptr = np.ascontiguousarray(rowptr, dtype=np.int64)
    col = np.ascontiguousarray(col, dtype=np.int32)
    val = np.ascontiguousarray(val, dtype=np.float64)
    return Matrix(lib().ko_matrix_from_csr(n, m, _p(rowptr), _p(col), _p(val)))


def from_dense(d):
    d = np.asarray(d, dtype=np.float64)
    rowptr = [0]
    col, val = [], []
    for r in d:
        nz = np.nonzero(r)[0]
        col.extend(nz.tolist())
        val.extend(r[nz].tolist())
        rowptr.append(len(col))
    return from_csr(d.shape[0], d.shape[1], rowptr, col or [0], val or [0.0])


def coeff_dim(n):
    return lib().ko_coeff_dim(n)


def ind2sub(n, k1, k2):
    return lib().ko_coeff_ind2sub(n, k1, k2)


def sub2ind(n, i):
    a, b = C.c_int64(), C.c_int64()
    lib().ko_coeff_sub2ind(n, i, a, b)
    return a.value, b.value


def ntheta(mat, cooccurrence):
    return coeff_dim(mat.m) if cooccurrence else mat.m + 1


def _lab(labels):
    return np.ascontiguousarray(labels, dtype=np.uint8)


def _cw(cw):
    return np.ascontiguousarray(cw, dtype=np.float64)


def linear_pdf(mat, theta, cooccurrence=False):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    out = np.zeros(mat.n)
    lib().ko_linear_pdf(mat.h, _p(theta), int(cooccurrence), _p(out))
    return out


def log_pdf(mat, theta, cooccurrence=False):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    out = np.zeros(mat.n)
    lib().ko_log_pdf(mat.h, _p(theta), int(cooccurrence), _p(out))
    return out


def gradient(mat, labels, theta, cw=(1.0, 1.0), lam=0.0, cooccurrence=False):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    g = np.zeros(len(theta))
    lab, cw = _lab(labels), _cw(cw)
    lib().ko_gradient(mat.h, _p(lab), _p(theta), len(theta), _p(cw), lam, int(cooccurrence), _p(g))
    return g


def loss(mat, labels, theta, cw=(1.0, 1.0), lam=0.0, cooccurrence=False):
    theta = np.ascontiguousarray(theta, dtype=np.float64)
    lab, cw = _lab(labels), _cw(cw)
    return lib().ko_loss(mat.h, _p(lab), _p(theta), len(theta), _p(cw), lam, int(cooccurrence))


def class_weights(labels):
    lab = _lab(labels)
    cw = np.zeros(2)
    lib().ko_class_weights(_p(lab), len(lab), _p(cw))
    return cw


def nlargest_abs(x, tie=TIE_GO118):
    x = np.array(x, dtype=np.float64)
    idx = np.zeros(len(x), dtype=np.int64)
    lib().ko_nlargest_abs(_p(x), _p(idx), len(x), tie)
    return x, idx


def select(mat, labels, cw, N, theta0, active_idx, active_theta, cooccurrence=False, tie=TIE_GO118,
           epsilon_lambda=0.0, prev_lambda=0.0):
    nt = ntheta(mat, cooccurrence)
    ai = np.ascontiguousarray(active_idx, dtype=np.int64)
    at = np.ascontiguousarray(active_theta, dtype=np.float64)
    mask = np.zeros(nt, dtype=np.uint8)
    g = np.zeros(nt)
    lam, c = C.c_double(), C.c_int64()
    lab, cw = _lab(labels), _cw(cw)
    ok = lib().ko_select(mat.h, _p(lab), _p(cw), int(cooccurrence), N, theta0, _p(ai), _p(at), len(ai), tie,
                         epsilon_lambda, prev_lambda, _p(mask), nt, lam, c, _p(g))
    return dict(ok=bool(ok), lam=lam.value, c=c.value, mask=mask.astype(bool), g=g)


def reduce(mat, sel):
    sel = np.ascontiguousarray(sel, dtype=np.int64)
    return Matrix(lib().ko_reduce(mat.h, _p(sel), len(sel)))


def step_size(mat, l2=0.0, step_factor=1.0):
    return lib().ko_step_size(mat.h, l2, step_factor)


def proxgrad(rmat, labels, theta, cw=(1.0, 1.0), lam=0.0, l2=0.0, step_factor=1.0, epsilon=0.0,
             epsilon_loss=1e-8, max_iter=100000, hook=None):
    theta = np.array(theta, dtype=np.float64)
    lab, cw = _lab(labels), _cw(cw)
    delta = C.c_double()
    hk = hook if hook is not None else HookState(float("nan"), float("nan"))
    it = lib().ko_proxgrad(rmat.h, _p(lab), _p(theta), _p(cw), lam, l2, step_factor, epsilon, epsilon_loss,
                           max_iter, C.byref(hk), delta)
    return theta, it, delta.value


class EstimatorState:
    """KmerLrEstimator state carried between leapfrog targets (warm start, kmerLr_estimator.go:263-267)."""

    def __init__(self, cap=4096):
        self.cap = cap
        self.idx = np.zeros(cap, dtype=np.int64)
        self.theta = np.zeros(cap + 1, dtype=np.float64)
        self.c = Estimator(0, self.idx.ctypes.data_as(C.POINTER(C.c_int64)),
                           self.theta.ctypes.data_as(C.POINTER(C.c_double)), cap,
                           HookState(float("nan"), float("nan")), 0.0)

    @property
    def active_idx(self):
        return self.idx[:self.c.n_active].copy()

    @property
    def active_theta(self):
        return self.theta[:self.c.n_active + 1].copy()


def estimate_loop(mat, labels, cw, N, est, cooccurrence=False, tie=TIE_GO118, epsilon_lambda=0.0, l2=0.0,
                  step_factor=1.0, epsilon=0.0, epsilon_loss=1e-8, max_iter=100000, max_epochs=0, path_cap=256):
    lab, cw = _lab(labels), _cw(cw)
    pl = np.zeros(path_cap)
    pi = np.zeros(path_cap, dtype=np.int64)
    ep = lib().ko_estimate_loop(mat.h, _p(lab), _p(cw), int(cooccurrence), N, tie, epsilon_lambda, l2, step_factor,
                                epsilon, epsilon_loss, max_iter, max_epochs, C.byref(est.c), _p(pl), _p(pi), path_cap)
    if ep < 0:
        raise RuntimeError("estimator state capacity exceeded")
    k = min(ep, path_cap)
    return dict(epochs=ep, lambdas=pl[:k].copy(), iters=pi[:k].copy())


def window_slots(length, W, step):
    return lib().ko_window_slots(length, W, step)


def score_windows(models, regions, W, step, threads=0):
    """models: list of dict(cfg, class_k, class_code, features, theta (members x F+1), summary)."""
    keep = []
    arr = (Model * len(models))()
    for i, md in enumerate(models):
        ck = np.ascontiguousarray(md["class_k"], dtype=np.int32)
        cc = np.ascontiguousarray(md["class_code"], dtype=np.uint64)
        ft = np.ascontiguousarray(md["features"], dtype=np.int32).reshape(-1, 2)
        th = np.ascontiguousarray(md["theta"], dtype=np.float64).reshape(-1, ft.shape[0] + 1)
        keep += [ck, cc, ft, th]
        arr[i] = Model(md["cfg"], len(ck), ck.ctypes.data_as(C.POINTER(C.c_int32)),
                       cc.ctypes.data_as(C.POINTER(C.c_uint64)), ft.shape[0],
                       ft.ctypes.data_as(C.POINTER(C.c_int32)), th.shape[0],
                       th.ctypes.data_as(C.POINTER(C.c_double)), SUMMARY[md.get("summary", "")])
    buf, off = flatten(regions) if not isinstance(regions, tuple) else regions
    total = sum(window_slots(int(off[i + 1] - off[i]), W, step) for i in range(len(off) - 1))
    out = np.zeros(max(total, 1))
    lib().ko_score_windows(arr, len(models), _p(buf), _p(off), len(off) - 1, W, step, _p(out), threads)
    return out[:total]


# ---------------------------------------------------------------------------------------------------
# data transforms (kmerLr_transform.go:40-252,584-629), restated with numpy on DENSIFIED rows -- the
# reference densifies too when an offset is present (:600-609).  Small inputs only.
# ---------------------------------------------------------------------------------------------------
def pair_dense(mat):
    """the pair features v_a v_b, a < b, as dense columns in CoeffIndex order (kmerLr_coefficients_index.go:26-54)"""
    X = mat.dense()
    n, m = X.shape
    cols = [X[:, a] * X[:, b] for a in range(m) for b in range(a + 1, m)]
    return np.stack(cols, axis=1) if cols else np.zeros((n, 0))


def fit_transform(mat, kind, cooccurrence=False):
    """TransformFull.Fit: returns (offset or None, scale or None), index 0 = bias; with cooccurrence the pair features
    follow the single features in CoeffIndex order (kmerLr_transform.go:59-252)"""
    X = mat.dense()
    if cooccurrence:
        X = np.concatenate([X, pair_dense(mat)], axis=1)
    n, m = X.shape
    kind = (kind or "none").lower()
    if kind in ("", "none"):
        return None, None
    if kind in ("standardizer", "variance-scaler"):
        offset = np.zeros(m + 1)
        scale = np.ones(m + 1)
        offset[1:] = X.sum(axis=0) / float(n)                       # :77-104
        for j in range(m):
            nz = X[:, j][X[:, j] != 0]
            sj = 0.0
            for v in nz:                                              # :106-127, samples in order
                sj += (v - offset[j + 1]) * (v - offset[j + 1])
            sj += float(n - len(nz)) * offset[j + 1] * offset[j + 1]  # :129-131 zero entries
            scale[j + 1] = 1.0 if sj == 0.0 else 1.0 / np.sqrt(sj / float(n - 1))
        offset[0] = 0.0
        return (offset, scale) if kind == "standardizer" else (None, scale)
    if kind == "max-abs-scaler":
        scale = np.ones(m + 1)
        with np.errstate(divide="ignore"):
            scale[1:] = 1.0 / np.abs(X).max(axis=0)                  # :161-205
        return None, scale
    if kind == "mean-scaler":
        scale = np.ones(m + 1)
        with np.errstate(divide="ignore"):
            scale[1:] = float(n) / X.sum(axis=0)                     # :207-252
        return None, scale
    raise ValueError("invalid data transform")


def _transformed_dense(mat, offset, scale, cooccurrence=False):
    X = mat.dense()
    if cooccurrence:
        X = np.concatenate([X, pair_dense(mat)], axis=1)
    if offset is not None:
        X = X - offset[None, 1:]
    if scale is not None:
        X = X * scale[None, 1:]
    return X


def _log_add0(x):
    return np.where(x > 0, x + np.log1p(np.exp(-np.abs(x))), np.log1p(np.exp(-np.abs(x))))


def transformed_loss(mat, labels, theta, offset, scale, cw=(1.0, 1.0), lam=0.0, cooccurrence=False):
    """Loss (kmerLr_logistic_regression.go:250-272) on Transform.Apply'd rows"""
    theta = np.asarray(theta, dtype=np.float64)
    z = theta[0] + _transformed_dense(mat, offset, scale, cooccurrence) @ theta[1:]
    lab = np.asarray(labels).astype(bool)
    r = 0.0
    for i in range(len(z)):                                           # serial over samples
        r -= cw[1] * (-_log_add0(-z[i])) if lab[i] else cw[0] * (-_log_add0(z[i]))
    r /= float(len(z))
    if lam == lam and lam != 0.0:
        r += lam * np.sum(np.abs(theta[1:mat.m + 1]))                  # the L1 loop bound is data[0].Dim() (:255,267)
    return float(r)


def transformed_gradient(mat, labels, theta, offset, scale, cw=(1.0, 1.0), lam=0.0, cooccurrence=False):
    """Gradient (kmerLr_logistic_regression.go:151-248) on Transform.Apply'd rows"""
    theta = np.asarray(theta, dtype=np.float64)
    X = _transformed_dense(mat, offset, scale, cooccurrence)
    z = theta[0] + X @ theta[1:]
    lab = np.asarray(labels).astype(bool)
    r = -_log_add0(-z)
    w = np.where(lab, cw[1] * (np.exp(r) - 1.0), cw[0] * np.exp(r)) / float(len(z))
    g = np.concatenate([[w.sum()], X.T @ w])
    if lam == lam and lam != 0.0:
        g[1:] += lam * np.sign(theta[1:])
    return g


def select_from_gradient(g, N, active_idx, active_theta, tie=TIE_GO118, epsilon_lambda=0.0, prev_lambda=0.0):
    """featureSelector.Select after its gradient call (kmerLr_feature_selection.go:88-134) + computeLambda (:179-192),
    for a gradient computed elsewhere (e.g. under a data transform)"""
    g = np.asarray(g, dtype=np.float64)
    nt = len(g)
    b = np.zeros(nt, dtype=bool)
    b[0] = True
    c = 0
    for i, th in zip(active_idx, active_theta):
        if th != 0.0:
            b[int(i)] = True
            c += 1
    vals, idx = nlargest_abs(g[1:], tie)
    top = min(nt - 1, 2 * N)
    ok = False
    for k in range(top):
        if c >= N:
            break
        if not b[idx[k] + 1] and vals[k] != 0.0:
            ok = True
            b[idx[k] + 1] = True
            c += 1
    for k in range(top):
        if c >= N:
            break
        if not b[idx[k] + 1]:
            b[idx[k] + 1] = True
            c += 1
    if c > N:
        ok = True
    lam = 0.0
    if N <= top:
        v = abs(vals[N - 1])
        a = np.abs(g[1:])
        below = a[a < v]
        lam = (v + (below.max() if len(below) else 0.0)) / 2.0
    ok = ok or (epsilon_lambda > 0.0 and abs(prev_lambda - lam) >= epsilon_lambda)
    return dict(ok=bool(ok), lam=float(lam), c=c, mask=b)


def apply_transform(rmat, offset, scale):
    """Transform.Apply (kmerLr_transform.go:584-629) on a (reduced) matrix: with an offset the rows turn dense"""
    X = rmat.dense()
    if offset is not None:
        X = X - offset[None, 1:]
    if scale is not None:
        X = X * scale[None, 1:]
    return from_dense(X)


def estimate_loop_transformed(mat, labels, cw, N, offset, scale, tie=TIE_GO118, epsilon=0.0, epsilon_loss=1e-8,
                              max_iter=100000, max_epochs=0):
    """estimate_loop (kmerLr_estimator.go:209-255) under a data transform: the selection gradient is taken under
    the transform (kmerLr_feature_selection.go:221-229), every reduced data set goes through Transform.Apply before
    the solver (kmerLr_estimator.go:148).  Returns dict(active_idx, theta, lambdas, iters)."""
    nt = mat.m + 1
    active_idx, active_theta, th0, l1reg = np.zeros(0, dtype=np.int64), np.zeros(0), 0.0, 0.0
    hk = HookState(float("nan"), float("nan"))
    lambdas, iters, theta, have = [], [], np.zeros(1), False
    epoch = 0
    while max_epochs == 0 or epoch < max_epochs:
        t = np.zeros(nt)
        t[0] = th0
        nz = active_theta != 0.0
        t[active_idx[nz]] = active_theta[nz]
        g = transformed_gradient(mat, labels, t, offset, scale, cw)
        r = select_from_gradient(g, N, active_idx, active_theta, tie, prev_lambda=l1reg)
        if not r["ok"] and have:
            break
        lam = r["lam"]
        l1reg = lam * mat.n
        sel = np.nonzero(r["mask"])[0].astype(np.int64)
        red = apply_transform(reduce(mat, sel), None if offset is None else offset[sel], None if scale is None else scale[sel])
        theta, it, _ = proxgrad(red, labels, t[sel], cw, lam=lam, epsilon=epsilon, epsilon_loss=epsilon_loss,
                                max_iter=max_iter, hook=hk)
        active_idx, active_theta, th0 = sel[1:], theta[1:], theta[0]
        lambdas.append(lam)
        iters.append(it)
        have = True
        epoch += 1
    return dict(active_idx=active_idx, theta=theta, lambdas=lambdas, iters=iters)


# ---------------------------------------------------------------------------------------------------
# estimate_coordinate (kmerLr_estimator_coordinate.go:31-139): numpy restatement, loops as in the Go code
# (the three theta slices de-aliased: theta0 = start of the outer iteration, theta0_ = before the sweep).
# The reference never calls this function and no test of it exists: PARITY UNPINNED -- the CUDA path is
# checked against this restatement only.
# ---------------------------------------------------------------------------------------------------
def _eval_stopping(xs, x1, eps):
    """eval_stopping (kmerLr_estimator_proximal.go:30-52) -> (stop, delta)"""
    if np.any(np.isnan(x1)):
        return True, float("nan")
    max_x = float(np.max(np.abs(x1))) if len(x1) else 0.0
    max_d = float(np.max(np.abs(x1 - xs))) if len(x1) else 0.0
    delta = max_d / max_x if max_x != 0.0 else max_d
    return ((max_x != 0.0 and max_d / max_x <= eps) or (max_x == 0.0 and max_d == 0.0)), delta


def coordinate(rmat, labels, theta, cw_hook=(1.0, 1.0), l1reg=0.0, l2reg=0.0, epsilon=0.0, epsilon_loss=0.0,
               max_iter=100, hook=None):
    """-> (theta, coordinate sweeps done, delta of the last eval_stopping).  hook = [loss_old, loss_new] list."""
    y = np.asarray(labels, dtype=bool)
    X = np.hstack([np.ones((rmat.n, 1)), rmat.dense()])          # convert_counts: column 0 = bias (:204)
    n, d = X.shape
    cw = class_weights(labels)                                   # compute_class_weights (:88)
    theta1 = np.array(theta, dtype=np.float64)
    hk = hook if hook is not None else [float("nan"), float("nan")]

    def run_hook(th):                                            # kmerLr_estimator_hook.go:46-99
        hk[0], hk[1] = hk[1], hk[0]
        if epsilon_loss != 0.0:
            hk[1] = loss(rmat, labels, th, cw_hook, l1reg / n)
            return abs(hk[0] - hk[1]) < epsilon_loss
        return False

    sweeps, delta, it = 0, 0.0, 0
    with np.errstate(all="ignore"):
        while it < max_iter:                                     # :96
            r = X @ theta1                                       # LinearPdf
            p = np.exp(-_log_add0(-r))
            w = p * (1.0 - p)
            z = np.where(y, r + (1.0 - p) / w, r + (0.0 - p) / w)
            w = w * np.where(y, cw[1], cw[0])
            theta0 = theta1.copy()                               # :112-115
            xy = X.T @ (w * z)                                   # :40-50
            xx = X.T @ (w[:, None] * X)
            norm = np.diag(xx).copy()
            while it < max_iter:                                 # :55
                theta0_ = theta1.copy()
                for j in range(d):                               # :57-72
                    t = xy[j] + norm[j] * theta1[j] - xx[j] @ theta1
                    if j > 0:
                        t = max(abs(t) - l1reg, 0.0) if t >= 0.0 else -max(abs(t) - l1reg, 0.0)
                    theta1[j] = t / (norm[j] + l2reg)
                sweeps += 1
                stop, delta = _eval_stopping(theta0_, theta1, epsilon)
                if stop or run_hook(theta1):
                    break
                it += 1
            stop, delta = _eval_stopping(theta0, theta1, epsilon)  # :117
            if stop or run_hook(theta1):
                break
            it += 1
    if np.any(np.isnan(theta1)):
        theta1[:] = np.nan
    return theta1, sweeps, delta


# ---------------------------------------------------------------------------------------------------
# on-disk formats (test infrastructure, like everything in this file): the reference's fmt verbs restated with
# Python's % operator -- C's printf and Go's fmt print the same digits for %e / %f / %d (both round the exact
# binary value half to even); only the names of the non-finite values differ (Go: NaN, +Inf, -Inf)
# ---------------------------------------------------------------------------------------------------
def go_exp(x):
    """Go's portable math.Exp (src/math/exp.go: exp + expmulti, the FreeBSD e_exp.c algorithm); Python floats are
    IEEE doubles and never fused, as the algorithm is written"""
    import math
    Ln2Hi, Ln2Lo, Log2e = 6.93147180369123816490e-01, 1.90821492927058770002e-10, 1.44269504088896338700e+00
    if x != x or x == math.inf:
        return x
    if x == -math.inf:
        return 0.0
    if x > 7.09782712893383973096e+02:
        return math.inf
    if x < -7.45133219101941108420e+02:
        return 0.0
    if -1.0 / (1 << 28) < x < 1.0 / (1 << 28):
        return 1.0 + x
    k = 0
    if x < 0:
        k = int(Log2e * x - 0.5)
    elif x > 0:
        k = int(Log2e * x + 0.5)
    hi = x - float(k) * Ln2Hi
    lo = float(k) * Ln2Lo
    P1, P2, P3 = 1.66666666666666657415e-01, -2.77777777770155933842e-03, 6.61375632143793436117e-05
    P4, P5 = -1.65339022054652515390e-06, 4.13813679705723846039e-08
    r = hi - lo
    t = r * r
    c = r - t * (P1 + t * (P2 + t * (P3 + t * (P4 + t * P5))))
    y = 1 - ((lo - (r * c) / (2 - c)) - hi)
    return math.ldexp(y, k)


def go_fmt(fmt, v):
    """one float64 through a Go verb like %e, %13e, %0.15f"""
    import math
    import re
    if v != v or math.isinf(v):
        w = re.match(r"%0?(\d*)", fmt).group(1)
        name = "NaN" if v != v else ("+Inf" if v > 0 else "-Inf")
        return name.rjust(int(w)) if w else name
    return fmt % v


def wiggle_text(regions, predictions, track_name, window_size, window_step):
    """saveWindowPredictionsWiggle (kmerLr_predict_genomic.go:37-60); regions = [(seqname, from), ...]"""
    out = ["track type=wiggle_0 name=%s\n" % track_name]
    for (name, frm), pred in zip(regions, predictions):
        out.append("fixedStep chrom=%s start=%d step=%d span=%d\n" % (name, frm + window_size // 2, window_step, window_step))
        for p in pred:
            out.append(go_fmt("%0.15f", go_exp(float(p))) + "\n")
    return "".join(out)


def export_kmers_text(names, dense_rows, transformed):
    """export_kmers (kmerLr_data.go:127-174): names joined by ',', then every row dense"""
    out = [",".join(names) + "\n"]
    for row in dense_rows:
        out.append(",".join(go_fmt("%e", float(v)) if transformed else "%d" % int(v) for v in row) + "\n")
    return "".join(out)


def path_text(estimator, lam, norm, theta):
    """KmerRegularizationPath.Export (kmerLr_estimator_path.go:41-73)"""
    out = []
    head = ""
    if len(estimator) > 0:
        head += "%9s " % "estimator"
    head += "%13s %13s %s\n" % ("lambda", "norm", "theta")
    out.append(head)
    for i in range(len(lam)):
        line = ""
        if len(estimator) > 0:
            line += "%9d " % estimator[i]
        line += "%s %s" % (go_fmt("%13e", lam[i]), go_fmt("%13e", norm[i]))
        for j, t in enumerate(theta[i]):
            line += (" " if j == 0 else ",") + go_fmt("%e", float(t))
        out.append(line + "\n")
    return "".join(out)


def format_duration(ns):
    """format_duration (kmerLr_estimator_trace.go:28-35) with the float64 arithmetic of time.Duration.Hours etc."""
    import math
    ns = int(ns)

    def split(unit):          # Duration.Hours(): integer part + remainder / unit (durations are not negative)
        return float(ns // unit) + float(ns % unit) / float(unit)
    hours, minutes, seconds = split(3600 * 10 ** 9), split(60 * 10 ** 9), split(10 ** 9)
    millis = float(ns // 10 ** 6)
    return "%02d:%02d:%02d:%02d.%03d" % (int(hours / 24), int(math.fmod(hours, 24)), int(math.fmod(minutes, 60)),
                                         int(math.fmod(seconds, 60)), int(math.fmod(millis, 1000)))


def trace_text(duration_ns, iteration, change, nonzero, lam, loss):
    """Trace.Export (kmerLr_estimator_trace.go:49-80)"""
    head = "%15s %9s %12s %8s" % ("duration", "iteration", "change", "nonzero")
    if len(lam) > 0:
        head += " %12s" % "lambda"
    if len(loss) > 0:
        head += " %12s" % "loss"
    out = [head + "\n"]
    for i in range(len(iteration)):
        line = "%15s %9d %s %8d" % (format_duration(duration_ns[i]), iteration[i], go_fmt("%12e", change[i]), nonzero[i])
        if len(lam) > 0:
            line += " " + go_fmt("%12e", lam[i])
        if len(loss) > 0:
            line += " " + go_fmt("%12e", loss[i])
        out.append(line + "\n")
    return "".join(out)
