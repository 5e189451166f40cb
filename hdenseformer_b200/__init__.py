"""hdenseformer_b200 -- B200-native H-DenseFormer 3D training / sliding-window hot path.

Drop-in surfaces (same names and call signatures as the reference repository):
    hdenseformer_b200.models.HDenseFormer : HDenseFormer, HDenseFormer_32, HDenseFormer_16
    hdenseformer_b200.loss.combine_loss   : CEPlusDice, DeepSuperloss
    hdenseformer_b200.loss.dice_loss      : DiceLoss, BinaryDiceLoss
    hdenseformer_b200.loss.cross_entropy  : CrossentropyLoss
    hdenseformer_b200.trainer             : train_step, DataParallelTrainer, inference_slidingwindow, cal_steps
"""
__version__ = "0.1.0"
