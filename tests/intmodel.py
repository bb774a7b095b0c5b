"""NumPy model of the integer digit arithmetic used by the CUDA projection path.

Because the tensor-core contraction is exact integer arithmetic with a single float32 rounding,
its output can be predicted bit for bit on the host.  This module is that prediction; the GPU
tests compare the kernels against it with ``assert_array_equal``.
"""
import numpy as np

FRAC = 30
EXP_MIN, EXP_MAX = -80, 100


def balanced_digits(x):
    """int array -> 4 int8 planes d with x = sum d[i] 256^i, d0..d2 in [-128,127]."""
    r = x.astype(np.int64)
    out = []
    for _ in range(3):
        lo = ((r & 0xFF) ^ 0x80) - 0x80          # sign-extended low byte
        out.append(lo.astype(np.int8))
        r = (r - lo) >> 8
    assert np.all(np.abs(r) <= 64)
    out.append(r.astype(np.int8))
    return np.stack(out)


def digits_to_int(d):
    return sum(d[i].astype(np.int64) << (8 * i) for i in range(4))


def row_exponents(data):
    """data (n_t, n_sel, 3) float32 -> e (3, n_t) with max|x| < 2^e."""
    m = np.abs(data).max(axis=1).T.astype(np.float32)             # (3, n_t)
    e = np.frexp(m)[1].astype(np.int32)
    e = np.where(m > 0, np.clip(e, EXP_MIN, EXP_MAX), EXP_MIN).astype(np.int32)
    return e


def digitize(data):
    """data (n_t, n_sel, 3) float32 -> (X (3, n_t, n_sel) int64, e (3, n_t))."""
    e = row_exponents(data)
    scale = np.ldexp(np.float32(1), FRAC - e).astype(np.float32)  # (3, n_t)
    x = np.rint(data.transpose(2, 0, 1) * scale[:, :, None]).astype(np.int64)
    return x, e


def phase_ints(k_vecs, mean_sel):
    """float32 k (n_k,3), r (n_sel,3) -> X (2 n_k, n_sel) int64, rows (cos k0, sin k0, cos k1, ...)."""
    theta = np.dot(k_vecs.astype(np.float32), mean_sel.astype(np.float32).T)        # float32 sgemm (= fma chain)
    c = np.cos(theta.astype(np.float64)).astype(np.float32)
    s = np.sin(theta.astype(np.float64)).astype(np.float32)
    x = np.empty((2 * k_vecs.shape[0], mean_sel.shape[0]), np.int64)
    x[0::2] = np.rint(c * np.float32(2.0 ** FRAC)).astype(np.int64)
    x[1::2] = np.rint(s * np.float32(2.0 ** FRAC)).astype(np.int64)
    return x


def project(xa, xb, e, max_pass=32768):
    """Exact model of psa_project: xa (rows, n_sel), xb (3, n_t, n_sel) ints, e (3, n_t) -> P (rows,3,n_t) f32."""
    rows, n_sel = xa.shape
    n_t = xb.shape[1]
    da, out = balanced_digits(xa), None
    for a0 in range(0, n_sel, max_pass):
        sl = slice(a0, min(a0 + max_pass, n_sel))
        p = np.empty((rows, 3, n_t), np.float32)
        for pol in range(3):
            db = balanced_digits(xb[pol])
            t = np.zeros((rows, n_t), np.int64)
            for cls in range(4):                                   # digit-pair class i + j = 3 + cls
                acc = np.zeros((rows, n_t), np.int64)
                for i in range(4):
                    j = 3 + cls - i
                    if 0 <= j < 4:
                        acc += da[i][:, sl].astype(np.int64) @ db[j][:, sl].astype(np.int64).T
                assert np.abs(acc).max() < 2 ** 31               # what the int32 TMEM accumulator must hold
                t += acc << (8 * cls)
            p[:, pol, :] = np.ldexp(t.astype(np.float32), (e[pol] - 36)[None, :].astype(np.int32))
        out = p if out is None else (out + p).astype(np.float32)
    return out
