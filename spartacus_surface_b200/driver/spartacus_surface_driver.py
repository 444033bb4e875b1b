"""Equivalent of the offline driver program (driver/spartacus_surface_driver.F90:20-300):
read namelists + input file, allocate the flux objects, prepare the LW emission, call radsurf,
scale and sum, optionally write the netCDF output file (radsurf_save.save_canopy_fluxes).

    python -m spartacus_surface_b200.driver.spartacus_surface_driver config.nam input.nc output.nc

is the counterpart of `bin/spartacus_surface config.nam input.nc output.nc`.

`solver` is the radsurf implementation; it defaults to the product library.
Tests pass the CPU oracle through the same code to obtain reference results.
"""
from ..radsurf_config import config_type
from ..radsurf_canopy_flux import canopy_flux_type
from ..radsurf_boundary_conds_out import boundary_conds_out_type
from ..radsurf_simple_spectrum import calc_simple_spectrum_lw
from .spartacus_surface_config import driver_config_type
from .spartacus_surface_read_input import read_input


class DriverResult:
    pass


def setup_case(namelist_path, input_path, radsurf_overrides=None, driver_overrides=None,
               legendre_gauss_init=None):
    """Namelists + input file -> inputs and zeroed outputs (driver:100-190)."""
    config = config_type().read(namelist_path)
    for k, v in (radsurf_overrides or {}).items():
        if not hasattr(config, k):
            raise ValueError(f"unknown &radsurf entry {k}")
        setattr(config, k, v)
    config.consolidate(legendre_gauss_init)
    driver_config = driver_config_type().read(namelist_path, **(driver_overrides or {}))
    r = DriverResult()
    r.config, r.driver_config = config, driver_config
    (r.canopy_props, r.sw_spectral_props, r.lw_spectral_props,
     r.top_flux_dn_sw, r.top_flux_dn_direct_sw, r.top_flux_dn_lw) = read_input(input_path, config, driver_config)
    allocate_outputs(r)
    if config.do_lw:
        r.lw_spectral_props.calc_monochromatic_emission(r.canopy_props)
        calc_simple_spectrum_lw(config, r.canopy_props, r.lw_spectral_props)
    return r


def allocate_outputs(r):
    config, cp = r.config, r.canopy_props
    ncol, ntotlay = cp.ncol, cp.ntotlay
    r.bc_out = boundary_conds_out_type().allocate(ncol, config.nsw if config.do_sw else 0,
                                                  config.nlw if config.do_lw else 0)
    r.sw_norm_dir = r.sw_norm_diff = r.lw_internal = r.lw_norm = None
    if config.do_sw:
        r.sw_norm_dir = canopy_flux_type().allocate(config, ncol, ntotlay, config.nsw, use_direct=True)
        r.sw_norm_diff = canopy_flux_type().allocate(config, ncol, ntotlay, config.nsw, use_direct=True)
    if config.do_lw:
        r.lw_internal = canopy_flux_type().allocate(config, ncol, ntotlay, config.nlw, use_direct=False)
        r.lw_norm = canopy_flux_type().allocate(config, ncol, ntotlay, config.nlw, use_direct=False)


def run_radsurf(r, solver=None, istartcol=None, iendcol=None):
    if solver is None:
        from ..radsurf_interface import radsurf as solver
    return solver(r.config, r.canopy_props, r.sw_spectral_props if r.config.do_sw else None,
                  r.lw_spectral_props if r.config.do_lw else None, r.bc_out, istartcol, iendcol,
                  r.sw_norm_dir, r.sw_norm_diff, r.lw_internal, r.lw_norm)


def scale_and_sum(r):
    """driver:250-261: dimensional fluxes from the normalised ones."""
    config, cp = r.config, r.canopy_props
    r.sw_flux = r.lw_flux = None
    if config.do_sw:
        r.sw_norm_dir.scale(cp.nlay, r.top_flux_dn_direct_sw)
        r.sw_norm_diff.scale(cp.nlay, r.top_flux_dn_sw - r.top_flux_dn_direct_sw)
        r.sw_flux = canopy_flux_type().allocate(config, cp.ncol, cp.ntotlay, config.nsw, use_direct=True)
        r.sw_flux.sum(r.sw_norm_dir, r.sw_norm_diff)
    if config.do_lw:
        r.lw_norm.scale(cp.nlay, r.top_flux_dn_lw)
        r.lw_flux = canopy_flux_type().allocate(config, cp.ncol, cp.ntotlay, config.nlw, use_direct=False)
        r.lw_flux.sum(r.lw_internal, r.lw_norm)


def run_case(namelist_path, input_path, solver=None, radsurf_overrides=None, driver_overrides=None,
             legendre_gauss_init=None, do_scale=True, output_path=None):
    r = setup_case(namelist_path, input_path, radsurf_overrides, driver_overrides, legendre_gauss_init)
    r.status = run_radsurf(r, solver)
    if do_scale:
        scale_and_sum(r)
    if output_path is not None:
        if not do_scale:
            raise ValueError("the output file holds the scaled and summed fluxes")
        from ..radsurf_save import save_canopy_fluxes
        save_canopy_fluxes(output_path, r.config, r.canopy_props, r.sw_flux, r.lw_flux)
    return r


def main(argv=None):
    import sys
    argv = sys.argv[1:] if argv is None else argv
    if len(argv) != 3:
        raise SystemExit("usage: spartacus_surface_driver config.nam input.nc output.nc")
    r = run_case(argv[0], argv[1], output_path=argv[2])
    if r.status:
        raise SystemExit(f"radsurf flagged {r.status} layer problems")


if __name__ == "__main__":
    main()
