from .HDenseFormer import HDenseFormer, HDenseFormer_16, HDenseFormer_32  # noqa: F401
from .HDenseFormer_2D import HDenseFormer_2D, HDenseFormer_2D_16, HDenseFormer_2D_32  # noqa: F401
