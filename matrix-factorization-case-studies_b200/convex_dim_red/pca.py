"""PCA / EOF reduction feeding the PCA -> AA / GPNH / k-means drivers
(``bin/run_jra55_pca_aa.py`` and friends use ``sklearn.decomposition.PCA``; the JRA-55 EOFs are
167 components of a 700 x 41 800 field, ``bin/run_jra55_pca_aa_wrapper.sh:39``).

Computed through the sample-space Gram matrix, which is the cheap side when
n_samples << n_features: the column means and the centring use this package's kernels,
K = Xc Xc' is the DMMA Gram kernel (2 T^2 d flops, the only large product), the principal
axes Xc' U / s come from the streaming reduce-over-samples pass and ``transform`` is the
reduce-over-features pass.  Only the T x T symmetric eigen-decomposition itself is delegated
to the vendor library (``torch.linalg.eigh``) -- it is O(T^3) setup work outside the
alternating-update path.

Attributes and sign convention follow ``sklearn.decomposition.PCA(svd_solver='full')``:
``components_`` rows have their largest-magnitude entry positive (``svd_flip`` with
``u_based_decision=False``).
"""

import numpy as np

from . import _backend as be


class PCA():
    """Principal component analysis of an (n_samples, n_features) matrix, n_samples <=
    a few thousand."""

    def __init__(self, n_components=None, whiten=False):
        self.n_components = n_components
        self.whiten = whiten

    # -- helpers --------------------------------------------------------------------
    @staticmethod
    def _centred_copy(X, mean=None):
        torch = be.require_cuda()
        lib = be.library()
        T, d = X.shape
        Xd = be._upload_padded(np.ascontiguousarray(X, dtype=np.float64))   # private: modified
        ldx = Xd.stride(0)
        if mean is None:
            mean_d = be.zeros(ldx)
            be.check(lib.cdr_column_moments(Xd.data_ptr(), ldx, T, d, mean_d.data_ptr(), None,
                                            be.stream_ptr()), 'cdr_column_moments')
        else:
            mean_d = be.zeros(ldx)
            mean_d[:d].copy_(torch.from_numpy(np.ascontiguousarray(mean, dtype=np.float64)))
        be.check(lib.cdr_center_columns(Xd.data_ptr(), ldx, T, d, mean_d.data_ptr(), -1.0,
                                        be.stream_ptr()), 'cdr_center_columns')
        return Xd, mean_d

    def fit(self, X, y=None):
        self._fit(np.asarray(X))
        return self

    def fit_transform(self, X, y=None):
        U, S = self._fit(np.asarray(X))
        scores = U * S
        if self.whiten:
            scores = U * np.sqrt(X.shape[0] - 1)
        return scores

    def _fit(self, X):
        torch = be.require_cuda()
        T, d = X.shape
        n = min(T, d) if self.n_components is None else int(self.n_components)
        if not 1 <= n <= min(T, d):
            raise ValueError('n_components=%r must be between 1 and min(n_samples, n_features)=%d'
                             % (self.n_components, min(T, d)))
        Xd, mean_d = self._centred_copy(X)
        K = be.gram(Xd, T, d)
        evals, evecs = torch.linalg.eigh(K[:, :T])            # ascending
        evals = torch.flip(evals, dims=[0]).clamp_min(0.0)
        evecs = torch.flip(evecs, dims=[1])
        S_all = torch.sqrt(evals)
        S = S_all[:n]
        # principal axes V' = diag(1/s) U' Xc, in chunks of <= 64 components
        comps = be.zeros(n, Xd.stride(0))
        ws = be.Workspace(T, d, min(n, be.MAX_COMPONENTS))
        for lo in range(0, n, be.MAX_COMPONENTS):
            hi = min(n, lo + be.MAX_COMPONENTS)
            L = be.zeros(hi - lo, be.round_up(T))
            inv_s = torch.where(S[lo:hi] > 0, 1.0 / S[lo:hi], torch.zeros_like(S[lo:hi]))
            L[:, :T].copy_((evecs[:, lo:hi] * inv_s).t())
            be.reduce_samples(L, L.stride(0), 1, Xd, T, d, hi - lo, comps[lo:hi], ws)
        components = be.to_host(comps, n, d)
        U = evecs[:, :n].cpu().numpy()
        # deterministic signs: the largest-magnitude entry of every component is positive
        idx = np.argmax(np.abs(components), axis=1)
        signs = np.sign(components[np.arange(n), idx])
        signs[signs == 0] = 1.0
        components *= signs[:, np.newaxis]
        U = U * signs[np.newaxis, :]

        S_host = S.cpu().numpy()
        ev_all = (evals / max(T - 1, 1)).cpu().numpy()
        self.n_components_ = n
        self.n_samples_, self.n_features_in_ = T, d
        self.mean_ = mean_d[:d].cpu().numpy()
        self.components_ = components
        self.singular_values_ = S_host
        self.explained_variance_ = ev_all[:n]
        total = ev_all[:min(T, d)].sum()
        self.explained_variance_ratio_ = self.explained_variance_ / total
        self.noise_variance_ = float(ev_all[n:min(T, d)].mean()) if n < min(T, d) else 0.0
        return U, S_host

    def transform(self, X):
        """(X - mean) components' through the reduce-over-features pass."""
        X = np.asarray(X)
        T, d = X.shape
        n = self.n_components_
        Xd, _ = self._centred_copy(X, self.mean_)
        out = np.empty((T, n))
        ws = be.Workspace(T, d, min(n, be.MAX_COMPONENTS))
        for lo in range(0, n, be.MAX_COMPONENTS):
            hi = min(n, lo + be.MAX_COMPONENTS)
            M = be.to_device_padded(self.components_[lo:hi])
            res = be.zeros(hi - lo, be.round_up(T))
            be.reduce_features(M, Xd, T, d, hi - lo, res, ws)
            out[:, lo:hi] = be.to_host(res, hi - lo, T).T
        if self.whiten:
            out /= np.sqrt(self.explained_variance_)
        return out

    def inverse_transform(self, scores):
        scores = np.asarray(scores, dtype=np.float64)
        if self.whiten:
            scores = scores * np.sqrt(self.explained_variance_)
        return scores.dot(self.components_) + self.mean_
