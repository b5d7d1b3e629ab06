"""The method surface of upstream's element classes, declared as tables.

Downstream Amira code (and upstream's tests) reach the fields of ``Gene`` / ``GeneMer`` / ``Read`` / ``Node`` /
``Edge`` through ``get_x`` / ``set_x`` / ``increment_x`` methods called by name, so the mirror classes must
offer exactly those names.  Rather than spelling out dozens of one-line methods, each mirror class lists
(method name, attribute) pairs and ``expose`` installs them."""
from __future__ import annotations

from operator import attrgetter


def _getter(attr):
    get = attrgetter(attr)
    return lambda self: get(self)


def _setter(attr, convert=None):
    def set_(self, value):
        value = convert(value) if convert else value
        setattr(self, attr, value)
        return value
    return set_


def _stepper(attr, step):
    """method() adds `step`; with step None, method(value) adds the value; both return the new total"""
    if step is None:
        def add(self, value):
            total = getattr(self, attr) + value
            setattr(self, attr, total)
            return total
        return add

    def bump(self):
        total = getattr(self, attr) + step
        setattr(self, attr, total)
        return total
    return bump


def expose(getters=(), setters=(), steppers=()):
    """class decorator: getters / setters are (method, attribute[, converter]); steppers (method, attribute, step)"""
    def install(cls):
        for method, attr in getters:
            setattr(cls, method, _getter(attr))
        for entry in setters:
            setattr(cls, entry[0], _setter(entry[1], entry[2] if len(entry) > 2 else None))
        for method, attr, step in steppers:
            setattr(cls, method, _stepper(attr, step))
        return cls
    return install
