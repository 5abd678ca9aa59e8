"""B200-native (sm_100a) implementation of the multi-task encoder-decoder hot path of
caumente/multi_task_breast_cancer: forward + backward of MTUNetPlusPlus / MTnnUNet / Multi_BTS_UNet and the
Dice + focal multi-task loss, behind the reference's nn.Module constructors."""

__version__ = "0.1.0"
