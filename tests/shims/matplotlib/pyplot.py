import numpy as np


class _Canvas:
    def __init__(self, fig):
        self.fig = fig

    def draw(self):
        return None

    def get_width_height(self):
        return self.fig.w, self.fig.h

    def buffer_rgba(self):
        return np.zeros((self.fig.h, self.fig.w, 4), dtype=np.uint8).tobytes()


class _Axes:
    def imshow(self, data, **kw):
        return object()


class Figure:
    def __init__(self, figsize=(4, 2)):
        self.w, self.h = int(figsize[0] * 10), int(figsize[1] * 10)
        self.canvas = _Canvas(self)


def subplots(figsize=(4, 2), **kw):
    return Figure(figsize), _Axes()


def colorbar(*a, **kw):
    return None


def close(*a, **kw):
    return None
