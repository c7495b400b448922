from . import S_diagrams, ST_diagrams, SU_diagrams, SV_diagrams     # noqa: F401
