"""Decoder solver (``nengo.solvers.LstsqL2`` with its Cholesky sub-solver; App. A.7)."""
import numpy as np
import scipy.linalg


class Solver:
    weights = False


class LstsqL2(Solver):
    def __init__(self, weights=False, reg=0.1):
        self.weights = bool(weights)
        self.reg = float(reg)

    def gram(self, A):
        """Regularised normal-equation factor, reusable for several targets."""
        m, n = A.shape
        sigma = self.reg * A.max()
        transpose = m < n
        G = A @ A.T if transpose else A.T @ A
        G[np.diag_indices_from(G)] += m * sigma ** 2
        return scipy.linalg.cho_factor(G, overwrite_a=True), transpose

    def solve(self, A, Y, factor=None):
        if factor is None:
            factor = self.gram(A)
        chol, transpose = factor
        b = Y if transpose else A.T @ Y
        x = scipy.linalg.cho_solve(chol, b)
        return A.T @ x if transpose else x

    def __call__(self, A, Y, rng=None):
        Y = np.asarray(Y, dtype=np.float64)
        X = self.solve(np.asarray(A, dtype=np.float64), Y.reshape(len(Y), -1))
        return X, {}
