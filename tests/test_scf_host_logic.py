"""Host-side logic of the SCF driver that needs no GPU: the DIIS-family coefficient problems built from the Gram
matrices the device step returns must be the ones the reference builds from the matrices themselves
(HartreeFock.jl:1273-1316), and the convergence norm is the reference's getErrorNrms (:1230-1237)."""
import numpy as np
import pytest

from quiqbox_b200 import hartreefock as hf


def _history(m, n, seed):
    rng = np.random.RandomState(seed)
    S = rng.uniform(-1, 1, (n, n)); S = S @ S.T + n * np.eye(n)
    X = hf.getOrthonormalization(S)
    Ds, Fs = [], []
    for _ in range(m):
        D = rng.uniform(-1, 1, (n, n)); Ds.append((D + D.T) / 2)
        F = rng.uniform(-1, 1, (n, n)); Fs.append((F + F.T) / 2)
    Es = list(rng.uniform(-3, -2, m))
    return S, X, Ds, Fs, Es


@pytest.mark.parametrize("method", ["DIIS", "EDIIS", "ADIIS"])
@pytest.mark.parametrize("m", [2, 5])
def test_gram_form_gives_the_same_coefficients(method, m):
    S, X, Ds, Fs, Es = _history(m, 7, 3 + m)
    Gdf = np.array([[np.vdot(Ds[i], Fs[j]) for j in range(m)] for i in range(m)])
    errs = [(X.T @ (F @ D @ S - S @ D @ F) @ X) for F, D in zip(Fs, Ds)]
    Gee = np.array([[np.vdot(errs[i], errs[j]) for j in range(m)] for i in range(m)])
    c_ref = hf._xdiis_coeff(method, Ds, Fs, Es, S, X)
    c_gram = hf._xdiis_coeff_gram(method, Gdf, Gee, Es)
    assert np.allclose(c_ref, c_gram, atol=1e-9), (c_ref, c_gram)
    assert abs(c_gram.sum() - 1.0) < 1e-9


def test_orthonormalisation_and_sign_convention():
    S, X, *_ = _history(1, 6, 1)
    assert np.allclose(X @ S @ X, np.eye(6), atol=1e-12)            # X = S^(-1/2), HartreeFock.jl:39-42
    F = np.diag(np.arange(6.0)) + 0.1
    C, e = hf.getC(X, F)
    assert np.all(C[0, :] >= 0) and np.allclose(C.T @ S @ C, np.eye(6), atol=1e-10)
    assert np.allclose(F @ C, S @ C * e, atol=1e-9)
