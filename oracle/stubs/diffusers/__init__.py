"""Minimal stand-in for the `diffusers` symbols the reference's hot-path files import.

TEST INFRASTRUCTURE ONLY.  It exists so the UNMODIFIED reference modules under /root/reference can be imported and run on
CPU in the build container to validate oracle/ and to generate tests/golden/ vectors.  Nothing in tokensgen_b200/ imports it.
The arithmetic restated here (FeedForward, CogVideoXDownsample3D/Upsample3D, DiagonalGaussianDistribution) follows the
published diffusers 0.31 semantics; it is the one part of the oracle that is "parity unpinned" (no diffusers checkout offline).
"""
__version__ = "0.31.0.dev0-stub"
