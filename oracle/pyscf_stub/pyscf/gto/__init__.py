from . import ft_ao  # noqa: F401
