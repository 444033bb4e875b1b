"""Offline-driver equivalents (driver/ of the reference): namelist overrides and
the netCDF input reader.  Outside the solver hot path; used to run the
reference's test fixtures through the library."""
from .spartacus_surface_config import driver_config_type
from .spartacus_surface_read_input import read_input
