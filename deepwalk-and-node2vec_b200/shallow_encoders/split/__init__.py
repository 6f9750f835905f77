"""Split algorithms of the downstream yardstick (reference package shallow_encoders/split): same names, same `_target_` paths."""
from shallow_encoders.split.core import (SplitAlgorithm, TrainTestRatioSplit, TrainValTestRatioSplit,  # noqa: F401
                                         TrainValTestStratifiedNSamplesSplit)
