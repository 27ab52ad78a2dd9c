"""Synapse parameter objects (only ``Lowpass`` is used by the reference)."""


class Synapse:
    pass


class Lowpass(Synapse):
    def __init__(self, tau):
        self.tau = float(tau)

    def __repr__(self):
        return f"Lowpass(tau={self.tau})"

    def __eq__(self, other):
        return isinstance(other, Lowpass) and other.tau == self.tau

    def __hash__(self):
        return hash(("Lowpass", self.tau))
