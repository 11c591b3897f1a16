"""Import shim (golden-vector generation only) for the absent ``exca`` package: just enough surface for the
reference's hot-path modules (``model.py``, ``pl_module.py``, ``modeling_utils/{losses,metrics,optimizers,models}``)
to import on CPU.  Nothing here computes anything."""
import contextlib
import inspect

import pydantic

from . import base, cachedict, helpers, utils  # noqa: F401

__version__ = "0.4.5"


class ConfDict(dict):
    pass


class _Infra(pydantic.BaseModel):
    model_config = pydantic.ConfigDict(extra="allow")
    folder: str | None = None
    cluster: str | None = None
    version: str = "0"
    mode: str = "cached"
    keep_in_ram: bool = False
    gpus_per_node: int = 1

    def apply(self, *args, **kwargs):
        if len(args) == 1 and callable(args[0]) and not kwargs:
            return args[0]
        return lambda fn: fn


class MapInfra(_Infra):
    pass


class TaskInfra(_Infra):
    pass
