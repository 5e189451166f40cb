from .data_loader import (Compose, DataGenerator, MRNormalize, PETandCTNormalize, ResidentVolumes, To_Tensor,
                          Trunc_and_Normalize, collate_batch, hdf5_reader)
from .transformer_3d import RandomCrop3D, RandomFlip3D, RandomTranslationRotationZoom3D

__all__ = ["Compose", "DataGenerator", "MRNormalize", "PETandCTNormalize", "ResidentVolumes", "To_Tensor",
           "Trunc_and_Normalize", "collate_batch", "hdf5_reader", "RandomCrop3D", "RandomFlip3D",
           "RandomTranslationRotationZoom3D"]
