!  mdb_c_binding.F90 -- ISO_C_BINDING interfaces of libmdpscu_b200.so (include/mdpscu_b200.h).
!
!  NOT COMPILED IN THIS REPOSITORY'S CI: the build image has no Fortran compiler (gfortran, nvfortran,
!  pgfortran, flang, ifort are all absent).  The same ABI is exercised from C (tests/test_abi_and_host.py
!  compiles the header as C99), from C++ (msmpscu_b200/csrc) and from Python/ctypes (every -m gpu test).
!  With any Fortran 2003 compiler:   <fc> -c mdb_c_binding.F90 mdb_shims.F90 ; link with -lmdpscu_b200
!
module MDB_C_BINDING
  use, intrinsic :: iso_c_binding
  implicit none

  integer(c_int), parameter :: MDB_ORDER_ORIGINAL = 0, MDB_ORDER_CELL = 1
  integer(c_int), parameter :: MDB_F_XP = 0, MDB_F_XP1 = 1, MDB_F_FP = 2, MDB_F_DIS = 3, MDB_F_EPOT = 4, MDB_F_EKIN = 5, &
                               MDB_F_DEN = 6, MDB_F_ITYP = 7, MDB_F_STATU = 8, MDB_F_GID = 9, MDB_F_GIDINV = 10,         &
                               MDB_F_IC = 11, MDB_F_KVOIS = 12, MDB_F_INDI = 13, MDB_F_NAC = 14, MDB_F_NAAC = 15,        &
                               MDB_F_IA1TH = 16
  integer(c_int), parameter :: MDB_POT_EAM = 0, MDB_POT_FS = 1
  integer(c_int), parameter :: MDB_FORCE = 1, MDB_VIRIAL = 2, MDB_EPOT = 4, MDB_DEN = 8

  interface
     integer(c_int) function mdb_device_count() bind(C, name="mdb_device_count")
       import :: c_int
     end function
     integer(c_int) function mdb_ctx_create(device_id, ctx) bind(C, name="mdb_ctx_create")
       import :: c_int, c_ptr
       integer(c_int), value :: device_id
       type(c_ptr)           :: ctx            ! mdb_ctx **
     end function
     subroutine mdb_ctx_destroy(ctx) bind(C, name="mdb_ctx_destroy")
       import :: c_ptr
       type(c_ptr), value :: ctx
     end subroutine
     integer(c_int) function mdb_sync(ctx) bind(C, name="mdb_sync")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mdb_box_set(ctx, nbox, napb, boxlow, boxsize, boxshape, ifpd, ngroup, mass) bind(C, name="mdb_box_set")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: nbox, napb, ngroup
       real(c_double)        :: boxlow(3), boxsize(3), boxshape(3,3), mass(*)
       integer(c_int)        :: ifpd(3)
     end function
     integer(c_int) function mdb_state_upload(ctx, field, host, order) bind(C, name="mdb_state_upload")
       import :: c_int, c_ptr
       type(c_ptr), value    :: ctx, host
       integer(c_int), value :: field, order
     end function
     integer(c_int) function mdb_state_download(ctx, field, host, order) bind(C, name="mdb_state_download")
       import :: c_int, c_ptr
       type(c_ptr), value    :: ctx, host
       integer(c_int), value :: field, order
     end function
     type(c_ptr) function mdb_devptr(ctx, field) bind(C, name="mdb_devptr")
       import :: c_int, c_ptr
       type(c_ptr), value    :: ctx
       integer(c_int), value :: field
     end function
     integer(c_int) function mdb_tables_set(ctx, pot_type, nkind, ntab, csi, potr, fpotr, potb, fpotb, nkind1, nembd, rhod, &
                                            fembd, dfembd, kpair, kembd, ru2max) bind(C, name="mdb_tables_set")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: pot_type, nkind, ntab, nkind1, nembd
       real(c_double), value :: csi, rhod, ru2max
       real(c_double)        :: potr(nkind,*), fpotr(nkind,*), potb(nkind,*), fpotb(nkind,*), fembd(nkind1,*), dfembd(nkind1,*)
       integer(c_int)        :: kpair(*), kembd(*)
     end function
     integer(c_int) function mdb_tables_clear(ctx) bind(C, name="mdb_tables_clear")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mdb_nlist_init(ctx, nb_rm, mxkvois) bind(C, name="mdb_nlist_init")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       real(c_double)        :: nb_rm(*)
       integer(c_int), value :: mxkvois
     end function
     integer(c_int) function mdb_nlist_build(ctx) bind(C, name="mdb_nlist_build")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mdb_nlist_copyout(ctx, kvois, indi, order) bind(C, name="mdb_nlist_copyout")
       import :: c_int, c_ptr
       type(c_ptr), value    :: ctx
       integer(c_int)        :: kvois(*), indi(*)
       integer(c_int), value :: order
     end function
     integer(c_int) function mdb_force(ctx, flags, vtensor) bind(C, name="mdb_force")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: flags
       real(c_double)        :: vtensor(3,3)
     end function
     integer(c_int) function mdb_predict(ctx, h) bind(C, name="mdb_predict")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       real(c_double), value :: h
     end function
     integer(c_int) function mdb_correct(ctx, h) bind(C, name="mdb_correct")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       real(c_double), value :: h
     end function
     integer(c_int) function mdb_ekin(ctx) bind(C, name="mdb_ekin")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mdb_epc_set(ctx, enable, te, alpha, cut, he) bind(C, name="mdb_epc_set")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       integer(c_int)     :: enable(*)
       real(c_double)     :: te(*), alpha(*), cut(*), he(*)
     end function
     integer(c_int) function mdb_epc_apply(ctx) bind(C, name="mdb_epc_apply")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mdb_run(ctx, itime0, nsteps, it0, nb_uptab, h) bind(C, name="mdb_run")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: itime0, nsteps, it0, nb_uptab
       real(c_double), value :: h
     end function
     !--- Cal_GlobalT_DEV / VelScaling_DEV / CheckTimestep_DEV (CommonGPU/MD_DiffScheme_GPU.F90:1042, :1390, :1214)
     integer(c_int) function mdb_global_t(ctx, curt) bind(C, name="mdb_global_t")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value :: ctx
       real(c_double)     :: curt
     end function
     integer(c_int) function mdb_vel_scaling(ctx, dt) bind(C, name="mdb_vel_scaling")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       real(c_double), value :: dt
     end function
     integer(c_int) function mdb_check_timestep(ctx, th, h2s2, dmx2, iflag) bind(C, name="mdb_check_timestep")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       real(c_double), value :: th, h2s2, dmx2
       integer(c_int)        :: iflag
     end function
     !--- Do_Steepest_Forsteps_DEV (CommonGPU/MD_SteepestScheme_GPU.F90:263-290)
     integer(c_int) function mdb_steepest(ctx, mxnumsteps, meth, alpha, maxdis, mindis, minepot, iflag, maxmove, delepot) &
                                          bind(C, name="mdb_steepest")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: mxnumsteps, meth
       real(c_double), value :: alpha, maxdis, mindis, minepot
       integer(c_int)        :: iflag
       real(c_double)        :: maxmove, delepot
     end function
     integer(c_int) function mdb_atomic_stress(ctx, d_avp) bind(C, name="mdb_atomic_stress")
       import :: c_int, c_ptr, c_devptr
       type(c_ptr), value    :: ctx
       type(c_devptr), value :: d_avp
     end function
     integer(c_int) function mdb_nlist_reorder_nearest(ctx, nearest) bind(C, name="mdb_nlist_reorder_nearest")
       import :: c_int, c_ptr
       type(c_ptr), value    :: ctx
       integer(c_int), value :: nearest
     end function
     integer(c_int) function mdb_thermalize(ctx, ti, seed, draw) bind(C, name="mdb_thermalize")
       import :: c_int, c_ptr, c_double, c_long_long
       type(c_ptr), value          :: ctx
       real(c_double), value       :: ti
       integer(c_long_long), value :: seed
       integer(c_int), value       :: draw
     end function
     integer(c_int) function mdb_lbfgs(ctx, mxnumsteps, msave, factr, pgtol, iflag, nfg, niter) bind(C, name="mdb_lbfgs")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: mxnumsteps, msave
       real(c_double), value :: factr, pgtol
       integer(c_int)        :: iflag, nfg, niter
     end function
     integer(c_int) function mdb_damping(ctx) bind(C, name="mdb_damping")
       import :: c_int, c_ptr
       type(c_ptr), value :: ctx
     end function
     integer(c_int) function mdb_dyndamp(ctx, mxnumsteps, h, minepot, iflag, delepot) bind(C, name="mdb_dyndamp")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: mxnumsteps
       real(c_double), value :: h, minepot
       integer(c_int)        :: iflag
       real(c_double)        :: delepot
     end function
     integer(c_int) function mdb_cg(ctx, mxnumsteps, meth, maxdis, mindis, minepot, iflag, delepot) bind(C, name="mdb_cg")
       import :: c_int, c_ptr, c_double
       type(c_ptr), value    :: ctx
       integer(c_int), value :: mxnumsteps, meth
       real(c_double), value :: maxdis, mindis, minepot
       integer(c_int)        :: iflag
       real(c_double)        :: delepot
     end function
  end interface

  !--- one context per process: the reference keeps its device state in module variables too
  type(c_ptr), save :: m_CTX = c_null_ptr

end module MDB_C_BINDING
