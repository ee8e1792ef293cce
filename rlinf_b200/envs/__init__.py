from .synthetic import SyntheticVectorEnv  # noqa: F401
