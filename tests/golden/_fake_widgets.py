"""Minimal stand-in for ipywidgets so that the LIVE reference's launch_smart_aligner / launch_deform_viewer can be driven headlessly by
make_golden.py: sliders clamp like ipywidgets' bounded floats, buttons remember their callbacks."""


class _Bounded:
    def __init__(self, description="", min=0, max=100, value=0, step=1):
        self.description, self.min, self.max, self.step = description, min, max, step
        self._observers = []
        self._value = min
        self.value = value

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, v):
        v = type(self)._cast(v)
        self._value = type(self)._cast(min(max(v, self.min), self.max))
        for cb in self._observers:
            cb({"new": self._value})

    def observe(self, cb, names=None):
        self._observers.append(cb)


class FloatSlider(_Bounded):
    _cast = staticmethod(float)


class IntSlider(_Bounded):
    _cast = staticmethod(int)


class Dropdown:
    def __init__(self, options=(), description="", value=None):
        self.options, self.description = list(options), description
        self._observers = []
        self._value = self.options[0] if value is None and self.options else value

    @property
    def value(self):
        return self._value

    @value.setter
    def value(self, v):
        old, self._value = self._value, v
        if v != old:
            for cb in self._observers:
                cb({"new": v, "old": old})

    def observe(self, cb, names=None):
        self._observers.append(cb)


class Button:
    instances = []

    def __init__(self, description="", button_style=""):
        self.description = description
        self._cb = None
        Button.instances.append(self)

    def on_click(self, cb):
        self._cb = cb

    def click(self):
        self._cb(self)


class Output:
    def __enter__(self):
        return self

    def __exit__(self, *a):
        return False


class _Box:
    def __init__(self, children=()):
        self.children = list(children)


VBox = HBox = _Box
