import contextlib

DISCRIMINATOR_FIELD = "name"


@contextlib.contextmanager
def environment_variables(**kwargs):
    yield
