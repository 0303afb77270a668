"""Host-only: the plan-cache key of a model callable (``mininf_b200/nn.py::_callable_key``).

The reference re-runs the model function on every loss evaluation
(``/root/reference/mininf/nn.py:223-225``), so a Python hyper-parameter changed between steps takes
effect at once. The engine caches the traced plan; the key therefore carries, by VALUE, every plain
constant the function captured, and survives re-created thin wrappers."""
import functools

from mininf_b200.nn import _callable_key

SCALE = 1.0


def _model_with_global():
    return SCALE * 2


def test_closure_constants_are_part_of_the_key_by_value():
    scale = 1.0

    def model():
        return scale

    first = _callable_key(model)
    assert _callable_key(model) == first
    scale = 2.5                      # rebinding the closure cell: same function object, new constant
    assert _callable_key(model) != first
    scale = 1.0
    assert _callable_key(model) == first


def test_defaults_and_named_globals_are_part_of_the_key():
    def model(width=3.0, *, depth=2):
        return width * depth

    first = _callable_key(model)
    model.__defaults__ = (4.0,)
    assert _callable_key(model) != first
    model.__defaults__ = (3.0,)
    assert _callable_key(model) == first
    model.__kwdefaults__ = {"depth": 5}
    assert _callable_key(model) != first

    global SCALE
    before = _callable_key(_model_with_global)
    SCALE = 3.0
    try:
        assert _callable_key(_model_with_global) != before
    finally:
        SCALE = 1.0
    assert _callable_key(_model_with_global) == before


def test_recreated_wrappers_keep_the_key_and_objects_enter_by_identity():
    class Holder:
        def model(self):
            return 1

    holder = Holder()
    assert _callable_key(holder.model) == _callable_key(holder.model)       # a new bound method per access
    assert _callable_key(holder.model) != _callable_key(Holder().model)

    payload = [1.0, "a"]
    big = object()

    def model(x, y, flag=None):
        return x

    a = functools.partial(model, payload, big, flag=True)
    b = functools.partial(model, payload, big, flag=True)
    assert _callable_key(a) == _callable_key(b)
    assert _callable_key(functools.partial(model, payload, big, flag=False)) != _callable_key(a)
    assert _callable_key(functools.partial(model, [1.0, "b"], big, flag=True)) != _callable_key(a)
    assert _callable_key(functools.partial(model, payload, object(), flag=True)) != _callable_key(a)


def test_a_lambda_recreated_per_step_keeps_the_key():
    payload = object()

    def make():
        return lambda: payload

    assert _callable_key(make()) == _callable_key(make())        # same code object, same captured object
    other = object()

    def make_other():
        return lambda: other

    assert _callable_key(make()) != _callable_key(make_other())
