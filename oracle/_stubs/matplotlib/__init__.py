"""Import stub (test infrastructure only): the reference imports matplotlib at module
scope (encoders/encoder_base.py:13-14) but never uses it on the hot path."""


def use(*a, **k):
    pass
