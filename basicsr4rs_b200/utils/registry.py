"""Name -> class registry with the lookup semantics of the reference's ``basicsr/utils/registry.py``
(:30-88): ``@ARCH_REGISTRY.register()`` decorator, duplicate names assert, ``get`` raises ``KeyError``.

When the archs of this package are dropped into the reference tree they import the
reference's own ``basicsr.utils.registry`` instead; this stand-alone twin exists so that the package is
importable without BasicSR installed (GPU box, tests, bench).
"""


class Registry:

    def __init__(self, name):
        self._name = name
        self._objects = {}

    def register(self, obj=None, suffix=None):
        def add(o):
            key = o.__name__ if suffix is None else f'{o.__name__}_{suffix}'
            assert key not in self._objects, f"An object named '{key}' was already registered in '{self._name}' registry!"
            self._objects[key] = o
            return o

        return add if obj is None else add(obj)

    def get(self, name, suffix='basicsr'):
        found = self._objects.get(name, self._objects.get(f'{name}_{suffix}'))
        if found is None:
            raise KeyError(f"No object named '{name}' found in '{self._name}' registry!")
        return found

    def __contains__(self, name):
        return name in self._objects

    def __iter__(self):
        return iter(self._objects.items())

    def keys(self):
        return self._objects.keys()


ARCH_REGISTRY = Registry('arch')
