"""Host-side, one-time construction of the LCT operators (product code).

The reference builds these in NumPy on the host as well
(/root/reference/models/tflct.py:32-79 calling utils/helper.py:35-69,72-125);
here they are produced directly in the compact forms the CUDA kernels consume:

* the sqrt(t) resampling operator ``mtx`` as CSR (about 2 non-zeros per row)
  instead of a dense M x M matrix (helper.py:35-69);
* the inverse light-cone filter as a Hermitian **half** spectrum
  ``(M+1, 2N, 2N)`` complex64 instead of two full ``(2M, 2N, 2N)`` float32
  arrays (tflct.py:55-65, helper.py:72-125).  The PSF has one voxel per
  (y, x) column, so its 3-D DFT is a per-plane 2-D FFT of an analytic
  time-axis phase; no ``(2M, 2N, 2N)`` array is ever materialised.
"""
from __future__ import annotations

import numpy as np
import scipy.fft as sfft

SNR = 1e-1            # tflct.py:42
LIGHT_SPEED = 3e8     # tflct.py:34


def resampling_csr(M: int):
    """CSR ``(rowptr, colidx, vals)`` of ``mtx`` (helper.py:35-69), float32.

    ``mtx[i, j]`` averages ``x**-0.5`` over ``x in (iM, (i+1)M]`` with
    ``ceil(sqrt(x)) - 1 == j``.  The values reproduce the reference's float32
    halving tree exactly: per output row only the columns of its band are
    carried through the ``0.5 * (even + odd)`` reduction.
    """
    if M < 2 or (M & (M - 1)):
        raise AssertionError("time size should be a power of 2")   # helper.py:40
    K = int(np.log2(M))
    rowptr = np.zeros(M + 1, dtype=np.int32)
    cols, vals = [], []
    for i in range(M):
        x = np.arange(i * M + 1, (i + 1) * M + 1).astype(np.float32)
        j = (np.ceil(np.sqrt(x)) - 1).astype(np.int64)
        w = (np.float32(1.0) / np.sqrt(x)).astype(np.float32)
        j0, j1 = int(j[0]), int(j[-1])
        band = np.where(j[None, :] == np.arange(j0, j1 + 1)[:, None], w[None, :], np.float32(0)).astype(np.float32)
        for _ in range(K):
            band = np.float32(0.5) * (band[:, 0::2] + band[:, 1::2])
        v = band[:, 0]
        keep = v != 0
        cols.append(np.arange(j0, j1 + 1)[keep])
        vals.append(v[keep])
        rowptr[i + 1] = rowptr[i] + int(keep.sum())
    return rowptr, np.concatenate(cols).astype(np.int32), np.concatenate(vals).astype(np.float32)


def csr_transpose(rowptr, colidx, vals, ncols):
    """CSR of the transpose (``mtxi = mtx.T``, helper.py:61)."""
    nrows = len(rowptr) - 1
    rows = np.repeat(np.arange(nrows, dtype=np.int32), np.diff(rowptr))
    order = np.lexsort((rows, colidx))
    t_rowptr = np.zeros(ncols + 1, dtype=np.int32)
    np.add.at(t_rowptr, colidx.astype(np.int64) + 1, 1)
    return np.cumsum(t_rowptr).astype(np.int32), rows[order].astype(np.int32), vals[order].astype(np.float32)


def csr_to_dense(rowptr, colidx, vals, ncols):
    out = np.zeros((len(rowptr) - 1, ncols), dtype=vals.dtype)
    rows = np.repeat(np.arange(len(rowptr) - 1), np.diff(rowptr))
    out[rows, colidx] = vals
    return out


def falloff(M: int, material: str):
    """Radiometric falloff ``gridz ** p`` (tflct.py:49-52,123-127), float32, length M."""
    g = np.arange(M, dtype=np.float32) / (M - 1)
    if material == "diffuse":
        return (g ** 4).astype(np.float32)
    if material == "specular":
        return (g ** 2).astype(np.float32)
    return np.ones(M, dtype=np.float32)


def psf_support(N: int, M: int, slope: float):
    """Support of the light-cone PSF (helper.py:72-125) without building the volume.

    Returns ``(z, y, x, val)``: voxel indices in the rolled/transposed
    ``(2M, 2N, 2N)`` layout and the common float32 value ``1/sqrt(count)``.
    The float32 arithmetic of helper.py:79-104 is reproduced operation by
    operation so the selected voxels (including exact ties) are the reference's.
    """
    g = np.arange(2 * N, dtype=np.float32)
    g = g / (2 * N - 1) * 2 - 1                                     # helper.py:79-80
    gz = np.arange(2 * M, dtype=np.float32)
    gz = gz / (2 * M - 1) * 2                                       # helper.py:87-88
    coef = np.float32((4 * slope) ** 2)                             # python float, cast when it meets float32
    g2 = g ** 2
    r2 = g2[:, None] + g2[None, :]                                  # helper.py:96 (sum is symmetric)
    t = (coef * r2).astype(np.float32)
    # |t - gz[k]| is minimised next to k ~ t*(2M-1)/2; examine a window around it
    k0 = np.floor(t.astype(np.float64) * (2 * M - 1) / 2.0).astype(np.int64)
    cand = np.clip(k0[..., None] + np.arange(-2, 4)[None, None, :], 0, 2 * M - 1)
    b = np.abs(t[..., None] - gz[cand]).astype(np.float32)          # helper.py:96-97
    c = b.min(axis=2, keepdims=True)                                # helper.py:100
    hit = np.abs(b - c) < np.float32(1e-8)                          # helper.py:103
    # the clip can repeat an index at the borders: count each voxel once
    i0, i1, ic = np.nonzero(hit)
    zyx = np.unique(np.stack([cand[i0, i1, ic], i0, i1], axis=1), axis=0)
    val = np.float32(1.0) / np.sqrt(np.float32(len(zyx)))           # helper.py:112
    z = zyx[:, 0]
    y = (zyx[:, 1] + N) % (2 * N)                                   # helper.py:115
    x = (zyx[:, 2] + N) % (2 * N)                                   # helper.py:116
    return z, y, x, val


def inverse_filter_half(N: int, M: int, slope: float, method: str = "lct", snr: float = SNR,
                        planes_per_chunk: int = 32):
    """Half spectrum ``(M+1, 2N, 2N)`` complex64 of the reference's ``invpsf``
    (tflct.py:57-65): ``conj(F)/(1/snr + |F|^2)`` for 'lct', ``conj(F)`` for
    'bp', with ``F = fftn(psf)`` evaluated in double precision.

    Only ``kt in [0, M]`` is kept: the data entering the filter are real, and
    the reference keeps the real part of the result (tflct.py:153), so the
    remaining planes are the Hermitian mirror of these.
    """
    z, y, x, val = psf_support(N, M, slope)
    out = np.empty((M + 1, 2 * N, 2 * N), dtype=np.complex64)
    flat = (y * (2 * N) + x).astype(np.int64)
    unique = len(np.unique(flat)) == len(flat)              # several z per (y, x) only on exact ties
    phase_table = np.exp((-2j * np.pi / (2 * M)) * np.arange(2 * M)) * float(val)      # built in double
    for k0 in range(0, M + 1, planes_per_chunk):
        kt = np.arange(k0, min(M + 1, k0 + planes_per_chunk))
        phase = phase_table[(kt[:, None] * z[None, :]) % (2 * M)]
        plane = np.zeros((len(kt), 4 * N * N), dtype=np.complex128)
        if unique:
            plane[:, flat] = phase
        else:
            np.add.at(plane, (slice(None), flat), phase)
        f = sfft.fft2(plane.reshape(len(kt), 2 * N, 2 * N), axes=(1, 2), workers=-1)
        if method == "lct":
            w = np.conj(f) / (1.0 / snr + f.real ** 2 + f.imag ** 2)        # tflct.py:60
        elif method == "bp":
            w = np.conj(f)                                                   # tflct.py:62
        else:
            raise ValueError(f"unknown method {method!r}")
        out[kt] = w.astype(np.complex64)
    return out


def laplacian_filter():
    """5x5x5 zero-mean Laplacian of Gaussian for ``method='bp'`` (helper.py:13-32)."""
    d = np.arange(-2, 3, dtype=np.float32)
    r2 = (d[:, None, None] ** 2 + d[None, :, None] ** 2 + d[None, None, :] ** 2)
    w = np.exp(-r2 / 2.0)
    w = w / np.sum(w)
    w1 = w * (r2 - 3.0)
    return (w1 - np.mean(w1)).astype(np.float32)


def slope_for(M: int, bin_len: float, wall_size: float) -> float:
    """``width / trange`` as the reference forms it (tflct.py:35-37,55)."""
    width = wall_size / 2.0
    bin_resolution = bin_len / LIGHT_SPEED
    trange = M * LIGHT_SPEED * bin_resolution
    return width / trange
