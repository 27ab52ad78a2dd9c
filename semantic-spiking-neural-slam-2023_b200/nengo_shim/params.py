"""``nengo.Default`` sentinel (used by the reference as a default *argument value*,
``sspslam/networks/binding.py:196,289``)."""


class _DefaultType:
    def __repr__(self):
        return "Default"

    def __bool__(self):
        return False


Default = _DefaultType()


def is_default(value):
    return value is Default
