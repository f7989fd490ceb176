"""Top-level `pca` module: the reference pickles its PCA object whole (`torch.save(pca)` -> weights/.../pca.pt), so
`torch.load` needs an importable `pca.PCA` with buffers named `mean_` and `components_` (reference pca.py; used at
longvgen/video_ipadapter/resampler.py:199-207,230-237 and longvgen/pipeline/pipeline_cogvideox_t2to.py:891-904).
Only two small dense products are on the inference path (transform / inverse_transform); they run in the object's own
dtype, fp32 in the shipped checkpoint, as plain library GEMMs."""
import torch
from torch import nn


class PCA(nn.Module):
    """Principal components by SVD of the centred data, with sklearn's deterministic sign convention (largest-|u| entry of
    every left singular vector made positive) so fits are reproducible against sklearn.decomposition.PCA."""

    def __init__(self, n_components=None):
        super().__init__()
        self.n_components = n_components

    @torch.no_grad()
    def fit(self, X: torch.Tensor) -> "PCA":
        keep = X.shape[1] if self.n_components is None else min(int(self.n_components), X.shape[1])
        mean = X.mean(dim=0, keepdim=True)
        U, _, Vh = torch.linalg.svd(X - mean, full_matrices=False)
        pivot = U.abs().argmax(dim=0)
        sign = torch.sign(U[pivot, torch.arange(U.shape[1], device=U.device)])
        self.register_buffer("mean_", mean)
        self.register_buffer("components_", (Vh * sign[:, None])[:keep].contiguous())
        return self

    def _fitted(self):
        if not hasattr(self, "components_"):
            raise RuntimeError("PCA.fit() must be called (or a fitted object loaded) first")

    def transform(self, X: torch.Tensor) -> torch.Tensor:
        self._fitted()
        return (X - self.mean_) @ self.components_.T

    def inverse_transform(self, Y: torch.Tensor) -> torch.Tensor:
        self._fitted()
        return Y @ self.components_ + self.mean_

    def fit_transform(self, X: torch.Tensor) -> torch.Tensor:
        return self.fit(X).transform(X)

    forward = transform
