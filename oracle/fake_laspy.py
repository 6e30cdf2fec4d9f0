"""A stand-in for the `laspy` package -- TEST INFRASTRUCTURE ONLY (laspy is not installed here and there is no network).

Just enough of the API the reference's data-preparation scripts touch (data_proc/1_get_windows_split.py:33-52,
2_preprocessing_filter_norm.py:38-55): `laspy.read(path)` returns an object whose dimensions are attributes holding numpy
arrays and whose `.points` can be indexed and assigned back (`las.points = las.points[np.where(mask)]` filters every
dimension). Tiles are registered per path with `register(path, dict_of_arrays)`.
"""
import numpy as np

_REGISTRY = {}


def register(path, fields):
    _REGISTRY[path] = {k: np.array(v, copy=True) for k, v in fields.items()}


class _Points:
    def __init__(self, fields):
        self.fields = fields

    def __getitem__(self, idx):
        return _Points({k: v[idx] for k, v in self.fields.items()})

    def __len__(self):
        return len(next(iter(self.fields.values())))


class FakeLas:
    def __init__(self, fields):
        object.__setattr__(self, "_fields", dict(fields))

    @property
    def points(self):
        return _Points(self._fields)

    @points.setter
    def points(self, pts):
        object.__setattr__(self, "_fields", dict(pts.fields))

    def __getattr__(self, name):
        try:
            return self._fields[name]
        except KeyError:
            raise AttributeError(name)

    def __len__(self):
        return len(self.points)


def read(path):
    return FakeLas(_REGISTRY[path])
