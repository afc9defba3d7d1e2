"""Drop-in namespace: ``from okrolearn_b200.okrolearn import *`` mirrors ``from okrolearn.okrolearn import *`` of the
reference (README.md:20-28) for the Conv/Dense train-step hot path."""
import numpy as np  # noqa: F401  (the reference's star-import also exposes np)

from ._lib import get_ctx, set_precision, FP32, TF32, BF16  # noqa: F401
from .tensor import Tensor, pinned_empty  # noqa: F401
from .layers import (DenseLayer, Conv1DLayer, Conv2DLayer, Conv3DLayer, ReLUActivationLayer, SigmoidActivationLayer,  # noqa: F401
                     TanhActivationLayer, SoftmaxActivationLayer, MaxPoolingLayer, AveragePoolingLayer, FlattenLayer,
                     UnflattenLayer, Unfold, Fold, ELUActivationLayer, SELUActivationLayer, LeakyReLUActivationLayer,
                     PReLUActivationLayer, SwishActivationLayer, GELUActivationLayer, SoftsignActivationLayer,
                     HardTanhActivationLayer, LinearActivationLayer, SoftminActivationLayer, LogSoftmaxActivationLayer,
                     BatchNormLayer, L1RegularizationLayer, L2RegularizationLayer, L3RegularizationLayer, DropoutLayer,
                     MaxUnpoolingLayer, AverageUnpoolingLayer, PaddingLayer, ZeroPaddingLayer, ReflectionPaddingLayer,
                     RNNLayer, LSTMLayer, GRULayer)
from .losses import CrossEntropyLoss, MSELoss, LossValue  # noqa: F401
from .schedulers import (LearningRateScheduler, StepLRScheduler, TimeBasedDecayScheduler, ExponentialDecayScheduler,  # noqa: F401
                         LinearDecayScheduler)
from .optimizers import AdamOptimizer, SGDOptimizer  # noqa: F401
from .network import NeuralNetwork  # noqa: F401
