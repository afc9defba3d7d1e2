"""okrolearn_b200: B200-native (sm_100a) implementation of okrolearn's Conv/Dense train-step hot path behind the
reference's Python Tensor / layer / NeuralNetwork API.  Host Python -> ctypes -> libokl_b200.so (include/okl.h)."""
__version__ = "0.1.0"
