"""Exception names the reference imports (``binding.py:6``, ``workingmemory.py:3``)."""


class NengoException(Exception):
    pass


class ValidationError(NengoException, ValueError):
    def __init__(self, msg, attr=None, obj=None):
        self.attr, self.obj = attr, obj
        super().__init__(msg)


class ObsoleteError(NengoException):
    def __init__(self, msg, since=None, url=None):
        super().__init__(msg)


class NetworkContextError(NengoException, RuntimeError):
    pass


class BuildError(NengoException, ValueError):
    pass


class SimulationError(NengoException, RuntimeError):
    pass


class SimulatorClosed(NengoException, RuntimeError):
    pass


class ReadonlyError(ValidationError):
    pass
