"""glove_tensorflow_b200 -- B200-native (sm_100a) drop-in for the training path of yxtay/glove-tensorflow.

Importing the package does not load CUDA; ``glove_tensorflow_b200.engine`` does, and fails loudly when
``libglove_b200.so`` is missing (there is no CPU fallback)."""
__version__ = "0.1.0"
