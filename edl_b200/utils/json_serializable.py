"""Reflection-based JSON (de)serialisation for state objects (reference:
python/edl/utils/json_serializable.py:26-61): every non-callable instance attribute is stored under
its name (leading underscores kept), nested ``SerializableBase`` values recurse, dict keys survive a
round trip as strings."""
import json


class SerializableBase:
    def to_json(self, filter_names=None):
        raise NotImplementedError

    def from_json(self, s):
        raise NotImplementedError


def _encode(v):
    if isinstance(v, SerializableBase):
        return {"__edl__": type(v).__name__, "v": v.to_dict()}
    if isinstance(v, dict):
        return {str(k): _encode(x) for k, x in v.items()}
    if isinstance(v, (list, tuple, set)):
        return [_encode(x) for x in v]
    return v


class Serializable(SerializableBase):
    """Subclasses may define ``_nested = {"attr": cls}`` / ``_nested_dict = {"attr": cls}`` so nested
    objects are rebuilt with the right type on load."""

    _nested = {}
    _nested_dict = {}

    def to_dict(self, filter_names=None):
        d = {}
        for k, v in self.__dict__.items():
            if filter_names and k in filter_names:
                continue
            if callable(v):
                continue
            d[k] = _encode(v)
        return d

    def to_json(self, filter_names=None):
        return json.dumps(self.to_dict(filter_names))

    def from_dict(self, d):
        for k, v in d.items():
            if isinstance(v, dict) and "__edl__" in v and k in self._nested:
                obj = self._nested[k]()
                obj.from_dict(v["v"])
                v = obj
            elif isinstance(v, dict) and k in self._nested_dict:
                out = {}
                for kk, vv in v.items():
                    obj = self._nested_dict[k]()
                    obj.from_dict(vv["v"] if isinstance(vv, dict) and "__edl__" in vv else vv)
                    out[kk] = obj
                v = out
            setattr(self, k, v)
        return self

    def from_json(self, s):
        if isinstance(s, (bytes, bytearray)):
            s = s.decode("utf-8")
        return self.from_dict(json.loads(s))

    def __eq__(self, other):
        return type(self) is type(other) and self.to_dict() == other.to_dict()

    def __ne__(self, other):
        return not self == other
