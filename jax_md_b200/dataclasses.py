"""Frozen dataclasses with `.set()` -- the host-side stand-in for the
reference's pytree dataclasses (jax_md/dataclasses.py:37-106): consumers use
`dataclasses.replace(obj, ...)`, `obj.set(...)`, `static_field()` and
`unpack`/`astuple` on them."""
import dataclasses as _dc

replace = _dc.replace
fields = _dc.fields
asdict = _dc.asdict
field = _dc.field


def static_field(**kw):
  return _dc.field(metadata={'static': True}, **kw)


def dataclass(cls):
  data = _dc.dataclass(frozen=True, eq=False)(cls)

  def _set(self, **kwargs):
    return _dc.replace(self, **kwargs)
  data.set = _set
  return data


def astuple(obj):
  return tuple(getattr(obj, f.name) for f in _dc.fields(obj))


unpack = astuple
