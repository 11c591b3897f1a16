import inspect


def validate_kwargs(func, kwargs):
    params = inspect.signature(func).parameters
    if any(p.kind == inspect.Parameter.VAR_KEYWORD for p in params.values()):
        return
    unknown = set(kwargs) - set(params)
    if unknown:
        raise ValueError(f"Unknown kwargs {unknown} for {func}")
