"""Split algorithms of the downstream yardstick under the reference's package path, so that YAML `_target_` strings such as
`shallow_encoders.split.TrainTestRatioSplit` (configs/sge_sg_cora.yaml:69-71) and `shallow_encoders.split.core.<name>` both resolve."""
from shallow_encoders.split import core as _core

__all__ = ['SplitAlgorithm', 'TrainTestRatioSplit', 'TrainValTestRatioSplit', 'TrainValTestStratifiedNSamplesSplit']
for _name in __all__:
    globals()[_name] = getattr(_core, _name)
