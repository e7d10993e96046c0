"""Re-export of the independent pure-PyTorch deformable unfold + matmul (lives in oracle/ so that
bench.py's cpu_baseline leg can time it too)."""
from oracle.torch_unfold import deform_conv_nd, deform_unfold  # noqa: F401
