! radsurf_interface_b200.F90 - drop-in replacement for the BODY of the reference
! module radsurf_interface (radsurf/radsurf_interface.F90:20-317).
!
! Same module name, same public procedure, same dummy-argument list and the
! same derived types: a host model (or the offline driver) links this object
! in place of radsurf_interface.o inside libradsurf.a together with
! libspartacus_b200.so, and every call of
!     call radsurf(config, canopy_props, sw_spectral_props, lw_spectral_props, &
!          &       bc_out, istartcol, iendcol, sw_norm_dir, sw_norm_diff,      &
!          &       lw_internal, lw_norm)
! is solved on the GPU.  The shim only takes c_loc() of every ALLOCATED member
! (unallocated -> c_null_ptr) and fills the bind(C) mirrors of
! include/spartacus_b200.h; arrays keep the reference layout (spectral index
! fastest, packed ragged layers, 1-based istartlay).
!
! NOTE: this file cannot be compiled in the build image (no Fortran compiler);
! it is kept in step with the header by hand.  INTEGRATION.md describes the
! link line and the two driver settings that matter on a GPU (do_parallel and
! nblocksize).
module radsurf_interface

  use, intrinsic :: iso_c_binding
  implicit none
  public :: radsurf

  integer, parameter :: SSB200_MAX_NSTREAM = 16

  type, bind(C) :: ssb200_legendre_gauss
    integer(c_int32_t) :: nstream, pad_
    real(c_double) :: mu(SSB200_MAX_NSTREAM), sin_ang(SSB200_MAX_NSTREAM), tan_ang(SSB200_MAX_NSTREAM)
    real(c_double) :: weight(SSB200_MAX_NSTREAM), hweight(SSB200_MAX_NSTREAM), vweight(SSB200_MAX_NSTREAM)
    real(c_double) :: vadjustment, vadjustment2
  end type ssb200_legendre_gauss

  type, bind(C) :: ssb200_config
    integer(c_int32_t) :: do_sw, do_lw, use_sw_direct_albedo
    integer(c_int32_t) :: do_vegetation, do_urban
    integer(c_int32_t) :: n_vegetation_region_forest, n_vegetation_region_urban
    integer(c_int32_t) :: nsw, nlw
    integer(c_int32_t) :: use_symmetric_vegetation_scale_forest
    integer(c_int32_t) :: use_symmetric_vegetation_scale_urban
    integer(c_int32_t) :: iverbose
    real(c_double) :: vegetation_isolation_factor_forest
    real(c_double) :: vegetation_isolation_factor_urban
    real(c_double) :: min_vegetation_fraction
    real(c_double) :: min_building_fraction
    type(ssb200_legendre_gauss) :: lg_sw_forest, lg_sw_urban, lg_lw_forest, lg_lw_urban
  end type ssb200_config

  type, bind(C) :: ssb200_canopy_properties
    integer(c_int32_t) :: ncol, ntotlay
    type(c_ptr) :: nlay, istartlay, i_representation
    type(c_ptr) :: cos_sza, dz
    type(c_ptr) :: building_fraction, building_scale
    type(c_ptr) :: veg_fraction, veg_scale, veg_ext
    type(c_ptr) :: veg_fsd, veg_contact_fraction
  end type ssb200_canopy_properties

  type, bind(C) :: ssb200_sw_spectral_properties
    integer(c_int32_t) :: nspec, pad_
    type(c_ptr) :: air_ext, air_ssa, veg_ssa, ground_albedo
    type(c_ptr) :: roof_albedo, wall_albedo, wall_specular_frac
    type(c_ptr) :: ground_albedo_dir, roof_albedo_dir
  end type ssb200_sw_spectral_properties

  type, bind(C) :: ssb200_lw_spectral_properties
    integer(c_int32_t) :: nspec, pad_
    type(c_ptr) :: air_ext, air_ssa, clear_air_planck
    type(c_ptr) :: veg_ssa, veg_planck, veg_air_planck
    type(c_ptr) :: ground_emissivity, ground_emission
    type(c_ptr) :: roof_emissivity, wall_emissivity
    type(c_ptr) :: roof_emission, wall_emission
  end type ssb200_lw_spectral_properties

  type, bind(C) :: ssb200_canopy_flux
    integer(c_int32_t) :: nspec, ncol, ntotlay, pad_
    type(c_ptr) :: ground_dn, ground_net, ground_vertical_diff, top_dn, top_net
    type(c_ptr) :: ground_dn_dir, top_dn_dir, ground_sunlit_frac
    type(c_ptr) :: roof_in, roof_net, wall_in, wall_net
    type(c_ptr) :: roof_in_dir, wall_in_dir
    type(c_ptr) :: roof_sunlit_frac, wall_sunlit_frac
    type(c_ptr) :: clear_air_abs, veg_abs, veg_air_abs
    type(c_ptr) :: veg_abs_dir, veg_sunlit_frac
    type(c_ptr) :: flux_dn_layer_top, flux_up_layer_top
    type(c_ptr) :: flux_dn_layer_base, flux_up_layer_base
    type(c_ptr) :: flux_dn_dir_layer_top, flux_dn_dir_layer_base
  end type ssb200_canopy_flux

  type, bind(C) :: ssb200_boundary_conds_out
    type(c_ptr) :: sw_albedo, sw_albedo_dir, lw_emissivity, lw_emission
  end type ssb200_boundary_conds_out

  interface
    function ssb200_radsurf(config, canopy_props, sw_spectral_props, lw_spectral_props, bc_out, &
         &  istartcol, iendcol, sw_norm_dir, sw_norm_diff, lw_internal, lw_norm) &
#ifdef SINGLE_PRECISION
         ! jprb = real32: the arrays behind the c_ptr members hold float; the library widens them on the
         ! device, solves in double precision and rounds the outputs (include/spartacus_b200.h)
         &  bind(C, name='ssb200_radsurf_sp') result(status)
#else
         &  bind(C, name='ssb200_radsurf') result(status)
#endif
      import
      type(ssb200_config),                 intent(in) :: config
      type(ssb200_canopy_properties),      intent(in) :: canopy_props
      type(c_ptr), value :: sw_spectral_props, lw_spectral_props
      type(ssb200_boundary_conds_out),     intent(in) :: bc_out
      integer(c_int32_t), value :: istartcol, iendcol
      type(c_ptr), value :: sw_norm_dir, sw_norm_diff, lw_internal, lw_norm
      integer(c_int) :: status
    end function ssb200_radsurf
    function ssb200_last_error() bind(C, name='ssb200_last_error') result(msg)
      import
      type(c_ptr) :: msg
    end function ssb200_last_error
  end interface

contains

  subroutine radsurf(config, canopy_props, sw_spectral_props, lw_spectral_props, &
       &             bc_out, istartcol, iendcol, sw_norm_dir, sw_norm_diff, &
       &             lw_internal, lw_norm)

    use parkind1,                       only : jpim, jprb
    use radiation_io,                   only : radiation_abort, nulerr
    use radsurf_config,                 only : config_type
    use radsurf_canopy_properties,      only : canopy_properties_type
    use radsurf_sw_spectral_properties, only : sw_spectral_properties_type
    use radsurf_lw_spectral_properties, only : lw_spectral_properties_type
    use radsurf_boundary_conds_out,     only : boundary_conds_out_type
    use radsurf_canopy_flux,            only : canopy_flux_type

    type(config_type),                 intent(in)    :: config
    type(canopy_properties_type),      intent(in), target :: canopy_props
    type(sw_spectral_properties_type), intent(in), target :: sw_spectral_props
    type(lw_spectral_properties_type), intent(in), target :: lw_spectral_props
    type(boundary_conds_out_type),     intent(inout), target :: bc_out
    integer(kind=jpim), optional,      intent(in)    :: istartcol, iendcol
    type(canopy_flux_type), intent(inout), optional, target :: sw_norm_dir, sw_norm_diff
    type(canopy_flux_type), intent(inout), optional, target :: lw_internal, lw_norm

    type(ssb200_config)                          :: c_config
    type(ssb200_canopy_properties)               :: c_canopy
    type(ssb200_sw_spectral_properties), target  :: c_sw
    type(ssb200_lw_spectral_properties), target  :: c_lw
    type(ssb200_boundary_conds_out)              :: c_bc
    type(ssb200_canopy_flux), target             :: c_f1, c_f2, c_f3, c_f4
    type(c_ptr) :: p_sw, p_lw, p_f1, p_f2, p_f3, p_f4
    integer(c_int32_t) :: icol1, icol2
    integer(c_int) :: status

    ! Column range: absent optionals select every column (reference :86-96).
    ! The reference driver's serial branch passes uninitialised values
    ! (driver/spartacus_surface_driver.F90:243-246), so clamp to 1..ncol.
    icol1 = 1
    icol2 = canopy_props%ncol
    if (present(istartcol)) icol1 = max(1, min(int(istartcol), canopy_props%ncol))
    if (present(iendcol)) then
      if (iendcol >= icol1 .and. iendcol <= canopy_props%ncol) icol2 = iendcol
    end if

    ! --- config_type (radsurf_config.F90:32-113) -------------------------------
    c_config%do_sw = merge(1, 0, config%do_sw)
    c_config%do_lw = merge(1, 0, config%do_lw)
    c_config%use_sw_direct_albedo = merge(1, 0, config%use_sw_direct_albedo)
    c_config%do_vegetation = merge(1, 0, config%do_vegetation)
    c_config%do_urban = merge(1, 0, config%do_urban)
    c_config%n_vegetation_region_forest = config%n_vegetation_region_forest
    c_config%n_vegetation_region_urban  = config%n_vegetation_region_urban
    c_config%nsw = config%nswinternal
    c_config%nlw = config%nlwinternal
    c_config%use_symmetric_vegetation_scale_forest = merge(1, 0, config%use_symmetric_vegetation_scale_forest)
    c_config%use_symmetric_vegetation_scale_urban  = merge(1, 0, config%use_symmetric_vegetation_scale_urban)
    c_config%iverbose = config%iverbose
    c_config%vegetation_isolation_factor_forest = config%vegetation_isolation_factor_forest
    c_config%vegetation_isolation_factor_urban  = config%vegetation_isolation_factor_urban
    c_config%min_vegetation_fraction = config%min_vegetation_fraction
    c_config%min_building_fraction   = config%min_building_fraction
    call copy_lg(config%lg_sw_forest, c_config%lg_sw_forest)
    call copy_lg(config%lg_sw_urban,  c_config%lg_sw_urban)
    call copy_lg(config%lg_lw_forest, c_config%lg_lw_forest)
    call copy_lg(config%lg_lw_urban,  c_config%lg_lw_urban)

    ! --- canopy_properties_type (radsurf_canopy_properties.F90:53-110) ---------
    c_canopy%ncol    = canopy_props%ncol
    c_canopy%ntotlay = canopy_props%ntotlay
    c_canopy%nlay              = c_loc(canopy_props%nlay)
    c_canopy%istartlay         = c_loc(canopy_props%istartlay)
    c_canopy%i_representation  = c_loc(canopy_props%i_representation)
    c_canopy%cos_sza           = loc1(canopy_props%cos_sza)
    c_canopy%dz                = loc1(canopy_props%dz)
    c_canopy%building_fraction = loc1(canopy_props%building_fraction)
    c_canopy%building_scale    = loc1(canopy_props%building_scale)
    c_canopy%veg_fraction      = loc1(canopy_props%veg_fraction)
    c_canopy%veg_scale         = loc1(canopy_props%veg_scale)
    c_canopy%veg_ext           = loc1(canopy_props%veg_ext)
    c_canopy%veg_fsd           = loc1(canopy_props%veg_fsd)
    c_canopy%veg_contact_fraction = loc1(canopy_props%veg_contact_fraction)

    p_sw = c_null_ptr
    p_f1 = c_null_ptr
    p_f2 = c_null_ptr
    if (config%do_sw) then
      c_sw%nspec = config%nswinternal
      c_sw%air_ext            = loc2(sw_spectral_props%air_ext)
      c_sw%air_ssa            = loc2(sw_spectral_props%air_ssa)
      c_sw%veg_ssa            = loc2(sw_spectral_props%veg_ssa)
      c_sw%ground_albedo      = loc2(sw_spectral_props%ground_albedo)
      c_sw%roof_albedo        = loc2(sw_spectral_props%roof_albedo)
      c_sw%wall_albedo        = loc2(sw_spectral_props%wall_albedo)
      c_sw%wall_specular_frac = loc2(sw_spectral_props%wall_specular_frac)
      c_sw%ground_albedo_dir  = loc2(sw_spectral_props%ground_albedo_dir)
      c_sw%roof_albedo_dir    = loc2(sw_spectral_props%roof_albedo_dir)
      p_sw = c_loc(c_sw)
      if (.not. (present(sw_norm_dir) .and. present(sw_norm_diff))) then
        call radiation_abort('radsurf: do_sw requires sw_norm_dir and sw_norm_diff')
      end if
      call copy_flux(sw_norm_dir,  c_f1)
      call copy_flux(sw_norm_diff, c_f2)
      p_f1 = c_loc(c_f1)
      p_f2 = c_loc(c_f2)
    end if

    p_lw = c_null_ptr
    p_f3 = c_null_ptr
    p_f4 = c_null_ptr
    if (config%do_lw) then
      c_lw%nspec = config%nlwinternal
      c_lw%air_ext           = loc2(lw_spectral_props%air_ext)
      c_lw%air_ssa           = loc2(lw_spectral_props%air_ssa)
      c_lw%clear_air_planck  = loc2(lw_spectral_props%clear_air_planck)
      c_lw%veg_ssa           = loc2(lw_spectral_props%veg_ssa)
      c_lw%veg_planck        = loc2(lw_spectral_props%veg_planck)
      c_lw%veg_air_planck    = loc2(lw_spectral_props%veg_air_planck)
      c_lw%ground_emissivity = loc2(lw_spectral_props%ground_emissivity)
      c_lw%ground_emission   = loc2(lw_spectral_props%ground_emission)
      c_lw%roof_emissivity   = loc2(lw_spectral_props%roof_emissivity)
      c_lw%wall_emissivity   = loc2(lw_spectral_props%wall_emissivity)
      c_lw%roof_emission     = loc2(lw_spectral_props%roof_emission)
      c_lw%wall_emission     = loc2(lw_spectral_props%wall_emission)
      p_lw = c_loc(c_lw)
      if (.not. (present(lw_internal) .and. present(lw_norm))) then
        call radiation_abort('radsurf: do_lw requires lw_internal and lw_norm')
      end if
      call copy_flux(lw_internal, c_f3)
      call copy_flux(lw_norm,     c_f4)
      p_f3 = c_loc(c_f3)
      p_f4 = c_loc(c_f4)
    end if

    c_bc%sw_albedo     = loc2(bc_out%sw_albedo)
    c_bc%sw_albedo_dir = loc2(bc_out%sw_albedo_dir)
    c_bc%lw_emissivity = loc2(bc_out%lw_emissivity)
    c_bc%lw_emission   = loc2(bc_out%lw_emission)

    status = ssb200_radsurf(c_config, c_canopy, p_sw, p_lw, c_bc, icol1, icol2, p_f1, p_f2, p_f3, p_f4)

    ! Error convention of the reference: message on unit nulerr, then abort
    ! (utilities/radiation_io.F90:35-54).  status > 0 counts layer problems
    ! with non-finite matrices, which the reference build would have trapped
    ! as a floating-point exception (Makefile_include.gfortran:28).
    if (status /= 0) then
      write(nulerr,'(a,i0,a)') '*** Error: ssb200_radsurf returned ', status, ': ' // c_message(ssb200_last_error())
      call radiation_abort('GPU radsurf failed')
    end if

  contains

    function loc1(a) result(p)
      real(kind=jprb), allocatable, target, intent(in) :: a(:)
      type(c_ptr) :: p
      p = c_null_ptr
      if (allocated(a)) p = c_loc(a)
    end function loc1

    function loc2(a) result(p)
      real(kind=jprb), allocatable, target, intent(in) :: a(:,:)
      type(c_ptr) :: p
      p = c_null_ptr
      if (allocated(a)) p = c_loc(a)
    end function loc2

    subroutine copy_lg(lg, c)
      use radtool_legendre_gauss, only : legendre_gauss_type
      type(legendre_gauss_type),   intent(in)  :: lg
      type(ssb200_legendre_gauss), intent(out) :: c
      integer :: n
      n = lg%nstream
      if (n > SSB200_MAX_NSTREAM) call radiation_abort('radsurf: more than 16 streams per hemisphere')
      c%nstream = n
      c%pad_ = 0
      c%mu = 0.0_c_double;      c%sin_ang = 0.0_c_double; c%tan_ang = 0.0_c_double
      c%weight = 0.0_c_double;  c%hweight = 0.0_c_double; c%vweight = 0.0_c_double
      c%mu(1:n)      = lg%mu(1:n)
      c%sin_ang(1:n) = lg%sin_ang(1:n)
      c%tan_ang(1:n) = lg%tan_ang(1:n)
      c%weight(1:n)  = lg%weight(1:n)
      c%hweight(1:n) = lg%hweight(1:n)
      c%vweight(1:n) = lg%vweight(1:n)
      c%vadjustment  = lg%vadjustment
      c%vadjustment2 = lg%vadjustment2
    end subroutine copy_lg

    subroutine copy_flux(f, c)
      type(canopy_flux_type), intent(inout), target :: f
      type(ssb200_canopy_flux), intent(out) :: c
      c%nspec = f%nspec;  c%ncol = f%ncol;  c%ntotlay = f%ntotlay;  c%pad_ = 0
      c%ground_dn            = loc2(f%ground_dn)
      c%ground_net           = loc2(f%ground_net)
      c%ground_vertical_diff = loc2(f%ground_vertical_diff)
      c%top_dn               = loc2(f%top_dn)
      c%top_net              = loc2(f%top_net)
      c%ground_dn_dir        = loc2(f%ground_dn_dir)
      c%top_dn_dir           = loc2(f%top_dn_dir)
      c%ground_sunlit_frac   = loc1(f%ground_sunlit_frac)
      c%roof_in              = loc2(f%roof_in)
      c%roof_net             = loc2(f%roof_net)
      c%wall_in              = loc2(f%wall_in)
      c%wall_net             = loc2(f%wall_net)
      c%roof_in_dir          = loc2(f%roof_in_dir)
      c%wall_in_dir          = loc2(f%wall_in_dir)
      c%roof_sunlit_frac     = loc1(f%roof_sunlit_frac)
      c%wall_sunlit_frac     = loc1(f%wall_sunlit_frac)
      c%clear_air_abs        = loc2(f%clear_air_abs)
      c%veg_abs              = loc2(f%veg_abs)
      c%veg_air_abs          = loc2(f%veg_air_abs)
      c%veg_abs_dir          = loc2(f%veg_abs_dir)
      c%veg_sunlit_frac      = loc1(f%veg_sunlit_frac)
      c%flux_dn_layer_top    = loc2(f%flux_dn_layer_top)
      c%flux_up_layer_top    = loc2(f%flux_up_layer_top)
      c%flux_dn_layer_base   = loc2(f%flux_dn_layer_base)
      c%flux_up_layer_base   = loc2(f%flux_up_layer_base)
      c%flux_dn_dir_layer_top  = loc2(f%flux_dn_dir_layer_top)
      c%flux_dn_dir_layer_base = loc2(f%flux_dn_dir_layer_base)
    end subroutine copy_flux

    function c_message(p) result(s)
      type(c_ptr), intent(in) :: p
      character(len=:), allocatable :: s
      character(kind=c_char), pointer :: chars(:)
      integer :: n
      s = ''
      if (.not. c_associated(p)) return
      call c_f_pointer(p, chars, [512])
      n = 0
      do while (n < 512)
        if (chars(n+1) == c_null_char) exit
        n = n + 1
      end do
      allocate(character(len=n) :: s)
      s = transfer(chars(1:n), s)
    end function c_message

  end subroutine radsurf

end module radsurf_interface
