"""Learning-rule parameter objects (``associativememory.py:31,41``)."""
from .params import Default, is_default
from .synapses import Lowpass, Synapse


def _syn(value, default):
    if is_default(value):
        return default
    if value is None or isinstance(value, Synapse):
        return value
    return Lowpass(float(value))


class LearningRuleType:
    modifies = None
    size_in = 0


class PES(LearningRuleType):
    modifies = "decoders"
    size_in = "post_state"

    def __init__(self, learning_rate=1e-4, pre_synapse=Default):
        self.learning_rate = float(learning_rate)
        self.pre_synapse = _syn(pre_synapse, Lowpass(0.005))


class Voja(LearningRuleType):
    modifies = "encoders"
    size_in = "scalar"

    def __init__(self, learning_rate=1e-2, post_synapse=Default):
        self.learning_rate = float(learning_rate)
        self.post_synapse = _syn(post_synapse, Lowpass(0.005))
