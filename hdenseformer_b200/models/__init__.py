from .HDenseFormer import HDenseFormer, HDenseFormer_16, HDenseFormer_32  # noqa: F401
