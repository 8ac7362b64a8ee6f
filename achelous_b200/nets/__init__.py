from .Achelous import Achelous, Achelous3T  # noqa: F401
