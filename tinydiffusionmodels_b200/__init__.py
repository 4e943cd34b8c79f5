"""tinydiffusionmodels_b200 — B200 (sm_100a) hot path of TinyDiffusionModels.

Host-side mirror of the reference's Python surface (src/mnist.py, src/shakespeare.py) over the
C-ABI CUDA library libtdm_b200.so.  See DESIGN.md.
"""
__version__ = "0.1.0"
