"""
shallow_encoders -- B200-native drop-in for the random-walk + skip-gram/negative-sampling hot path of
Robotmurlock/Deepwalk-and-Node2vec.  Same import paths and call signatures as the reference package for that path
(graph.random_walk_generator, graph.datasets, word2vec.model / loss / trainer / utils.sampling,
word2vec.dataloader.torch_dataset / registry, config_parser); all arithmetic runs in hand-written sm_100a CUDA
kernels behind the C ABI declared in include/se_b200.h (loaded by `shallow_encoders._native`).
There is no CPU fallback: compute entry points raise when the native library or a CUDA device is missing.
"""
__version__ = '0.1.0'
