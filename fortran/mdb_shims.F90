!  mdb_shims.F90 -- drop-in bodies for the reference procedures on the hot path.  Each keeps the reference's
!  name and dummy arguments, so MD_ForceClass_Register_GPU.F90 (Register_ForceClass), MD_Method_GenericMD_GPU.F90
!  (For_One_Step) and the application shell link unchanged; the CUDA-Fortran module bodies they replace are listed.
!  NOT COMPILED HERE (no Fortran compiler in the build image); see INTEGRATION.md.
!
module MD_EAM_Force_Table_GPU            ! replaces MDLIB/sor/CommonGPU/MD_EAM_ForceTable_GPU.F90
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MD_TYPEDEF_ForceTable
  use MDB_C_BINDING
  implicit none
contains
  subroutine INITIALIZE_EAM_Force_Table_DEV(SimBox, CtrlParam, FTable, RelaseTable, MULTIBOX)   ! :192-278
    type(SimMDBox),     intent(inout)::SimBox
    type(SimMDCtrl),    intent(in)   ::CtrlParam
    type(MDForceTable), intent(in)   ::FTable
    integer,            optional     ::RelaseTable, MULTIBOX
    integer(c_int)::ERR
      ERR = mdb_tables_set(m_CTX, MDB_POT_EAM, size(FTable%POTR,1), size(FTable%POTR,2), FTable%CSI,            &
                           FTable%POTR, FTable%FPOTR, FTable%POTB, FTable%FPOTB,                                 &
                           size(FTable%FEMBD,1), size(FTable%FEMBD,2), FTable%RHOD, FTable%FEMBD, FTable%DFEMBD, &
                           FTable%KPAIR, FTable%KEMBD, maxval(CtrlParam%RU*CtrlParam%RU))
      if(ERR .lt. 0) stop "MDPSCU Error: mdb_tables_set failed"
  end subroutine
  subroutine CALFORCE_EAM_Force_Table2A_DEV(SimBox, CtrlParam)                                   ! :950-1010
    type(SimMDBox),  intent(inout)::SimBox
    type(SimMDCtrl), intent(in)   ::CtrlParam
    real(c_double)::VT(3,3)
      if(mdb_force(m_CTX, MDB_FORCE, VT) .lt. 0) stop "MDPSCU Error: mdb_force failed"
  end subroutine
  subroutine CALPTENSOR_EAM_Force_Table2A_DEV(SimBox, CtrlParam)                                 ! :1366-1430
    type(SimMDBox),  intent(inout)::SimBox
    type(SimMDCtrl), intent(in)   ::CtrlParam
      if(mdb_force(m_CTX, ior(MDB_FORCE, MDB_VIRIAL), SimBox%VTENSOR) .lt. 0) stop "MDPSCU Error: mdb_force failed"
  end subroutine
  subroutine UpdateEPOT_EAM_Force_Table2A_DEV(SimBox, CtrlParam)                                 ! :1710-1745
    type(SimMDBox),  intent(inout)::SimBox
    type(SimMDCtrl), intent(in)   ::CtrlParam
    real(c_double)::VT(3,3)
      if(mdb_force(m_CTX, MDB_EPOT, VT) .lt. 0) stop "MDPSCU Error: mdb_force failed"
  end subroutine
  subroutine CALDEN_EAM_Force_Table2A_DEV(SimBox, CtrlParam)                                     ! :891-945
    type(SimMDBox),  intent(inout)::SimBox
    type(SimMDCtrl), intent(in)   ::CtrlParam
    real(c_double)::VT(3,3)
      if(mdb_force(m_CTX, MDB_DEN, VT) .lt. 0) stop "MDPSCU Error: mdb_force failed"
  end subroutine
  subroutine Cal_EAM_AtomicStressTensor_DEV(IDEV, dAVP)                                          ! :1973-1990 (pCalAVStress)
    integer, intent(in)::IDEV
    real(KINDDF), device, dimension(:,:), intent(out)::dAVP      ! the caller's device array, untouched CUDA-Fortran analysis code
      if(mdb_atomic_stress(m_CTX, c_devloc(dAVP)) .lt. 0) stop "MDPSCU Error: mdb_atomic_stress failed"
  end subroutine
  subroutine Clear_EAM_Force_Table_DEV()                                                         ! :343-365
    integer(c_int)::ERR
      ERR = mdb_tables_clear(m_CTX)
  end subroutine
end module MD_EAM_Force_Table_GPU

module MD_NeighborsList_GPU              ! replaces MDLIB/sor/CommonGPU/MD_NeighborsList_GPU.F90 (entry points :116-202)
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MDB_C_BINDING
  implicit none
contains
  subroutine Initialize_NeighboreList_DEV(SimBox, CtrlParam)                                     ! :241-396
    type(SimMDBox) ::SimBox
    type(SimMDCtrl)::CtrlParam
      if(mdb_nlist_init(m_CTX, CtrlParam%NB_RM, CtrlParam%NB_MXNBS) .lt. 0) stop "MDPSCU Error: mdb_nlist_init failed"
  end subroutine
  subroutine Cal_NeighBoreList_DEV(SimBox, CtrlParam)                                            ! :1347-1732
    type(SimMDBox) ::SimBox
    type(SimMDCtrl)::CtrlParam
    integer(c_int)::NOUT
      NOUT = mdb_nlist_build(m_CTX)
      if(NOUT .lt. 0) stop "MDPSCU Error: mdb_nlist_build failed"
      if(NOUT .gt. 0) write(*,fmt="(A, I8, A)") " MDPSCU Warning: there are ",NOUT," atoms out of box found in MD_NeighborsList."
  end subroutine
  subroutine Reorder_NeighBoreList_Nearest_Dev(Nearest)                                          ! :2016-2066
    integer::Nearest
      if(mdb_nlist_reorder_nearest(m_CTX, Nearest) .lt. 0) then
         write(*,*) "MDPSCU Error: the number of required neighbors (NEAREST) in Reorder_NeighBoreList_Nearest_DEV is larger than the permitted value"
         stop
      end if
  end subroutine
end module MD_NeighborsList_GPU

module MD_DiffScheme_GPU                 ! replaces MDLIB/sor/CommonGPU/MD_DiffScheme_GPU.F90 (:604, :821, :1000)
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MDB_C_BINDING
  implicit none
contains
  subroutine Do_DynDamp_Forsteps_DEV(SimBox, CtrlParam, ForceClass, MXNUMSTEPS)                  ! :1809-1860
    use MD_Forceclass_Register_GPU
    type(SimMDBox), dimension(:)      ::SimBox
    type(SimMDCtrl),       intent(in) ::CtrlParam
    type(MDForceClassGPU), intent(in) ::ForceClass
    integer,               intent(in) ::MXNUMSTEPS
    integer(c_int)::IFLAG
    real(c_double)::DELEPOT
      if(mdb_dyndamp(m_CTX, MXNUMSTEPS, CtrlParam%H, CtrlParam%STEEPEST_MiDelE*CP_EV2ERG, IFLAG, DELEPOT) .lt. 0) &
         stop "MDPSCU Error: mdb_dyndamp failed"
  end subroutine
  subroutine Predictor_DEV(ITIME, SimBox, CtrlParam)
    integer, intent(in)::ITIME
    type(SimMDBox)     ::SimBox
    type(SimMDCtrl)    ::CtrlParam
      if(ITIME .ge. CtrlParam%DAMPTIME0 .and. ITIME .le. CtrlParam%DAMPTIME0 + CtrlParam%DAMPTIME1-1) then   ! :611-617
         if(mdb_damping(m_CTX) .lt. 0) stop "MDPSCU Error: mdb_damping failed"
      end if
      if(mdb_predict(m_CTX, CtrlParam%H) .lt. 0) stop "MDPSCU Error: mdb_predict failed"
  end subroutine
  subroutine Correction_DEV(ITIME, SimBox, CtrlParam)
    integer, intent(in)::ITIME
    type(SimMDBox)     ::SimBox
    type(SimMDCtrl)    ::CtrlParam
      if(mdb_correct(m_CTX, CtrlParam%H) .lt. 0) stop "MDPSCU Error: mdb_correct failed"
  end subroutine
  subroutine CalEKin_DEV(SimBox, CtrlParam)
    type(SimMDBox) ::SimBox
    type(SimMDCtrl)::CtrlParam
      if(mdb_ekin(m_CTX) .lt. 0) stop "MDPSCU Error: mdb_ekin failed"
  end subroutine
  subroutine Cal_GlobalT_DEV(SimBox, CtrlParam, CURT)                                             ! :1042-1064
    type(SimMDBox) ::SimBox
    type(SimMDCtrl)::CtrlParam
    real(KINDDF)   ::CURT
      if(mdb_global_t(m_CTX, CURT) .lt. 0) stop "MDPSCU Error: mdb_global_t failed"
  end subroutine
  subroutine VelScaling_DEV(SimBox, CtrlParam, DT)                                                ! :1390-1446
    type(SimMDBox) ::SimBox
    type(SimMDCtrl)::CtrlParam
    real(KINDDF)   ::DT
      if(mdb_vel_scaling(m_CTX, DT) .lt. 0) then
         write(*,fmt="(A)") " MDPSCU Error: current temperature of the system is zero in scaling velocity"
         stop
      end if
  end subroutine
  subroutine Thermalizing_MC_DEV(SimBox, CtrlParam, TI)                                           ! :1746-1805
    type(SimMDBox), dimension(:)::SimBox
    type(SimMDCtrl)             ::CtrlParam
    real(KINDDF)                ::TI
    integer, save               ::DRAW = 0
      if(mdb_thermalize(m_CTX, TI, int(CtrlParam%SEED(1), c_long_long), DRAW) .lt. 0) stop "MDPSCU Error: mdb_thermalize failed"
      DRAW = DRAW + 1
  end subroutine
  subroutine CheckTimestep_DEV(ITIME, SimBox, CtrlParam, TH, H2S2, DMX2, IFLAG)                   ! :1214-1258
    integer,         intent(in)::ITIME
    type(SimMDBox),  intent(in)::SimBox
    type(SimMDCtrl)            ::CtrlParam
    real(KINDDF),    intent(in)::TH, H2S2, DMX2
    integer                    ::IFLAG
      if(mdb_check_timestep(m_CTX, TH, H2S2, DMX2, IFLAG) .lt. 0) stop "MDPSCU Error: mdb_check_timestep failed"
  end subroutine
end module MD_DiffScheme_GPU

module MD_SteepestScheme_GPU             ! replaces MDLIB/sor/CommonGPU/MD_SteepestScheme_GPU.F90 (:263-290)
  use MD_CONSTANTS
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MDB_C_BINDING
  implicit none
contains
  subroutine Do_Steepest_Forsteps_DEV(SimBox, CtrlParam, ForceClass, MXNUMSTEPS, METH)
    use MD_Forceclass_Register_GPU
    type(SimMDBox),dimension(:)       :: SimBox
    type(SimMDCtrl),       intent(in) :: CtrlParam
    type(MDForceClassGPU), intent(in) :: ForceClass
    integer,               intent(in) :: MXNUMSTEPS, METH
    integer(c_int)::IFLAG
    real(c_double)::MAXMOVE, DELEPOT
      if(mdb_steepest(m_CTX, MXNUMSTEPS, METH, CtrlParam%STEEPEST_Alpha, CtrlParam%STEEPEST_MxStep*SimBox(1)%RR,   &
                      CtrlParam%STEEPEST_MiStep*SimBox(1)%RR, CtrlParam%STEEPEST_MiDelE*CP_EV2ERG,                 &
                      IFLAG, MAXMOVE, DELEPOT) .lt. 0) stop "MDPSCU Error: mdb_steepest failed"
      if(IFLAG .eq. 0) then
         write(*,fmt="(A, I8)")      " MDPSCU WARNING: steepest finished after max steps: ", MXNUMSTEPS
         write(*,fmt="(A, 1PE13.4)") "                 with the max movement (LU) of atoms:", MAXMOVE/SimBox(1)%RR
      else
         write(*,fmt="(A, I8)")      " MDPSCU Message: steepest finished after steps: ", IFLAG
         write(*,fmt="(A, 1PE13.4)") "                 with max energy uncertainty(ev): ", DELEPOT*CP_ERG2EV
      end if
  end subroutine
end module MD_SteepestScheme_GPU

!--- CommonGPU/MD_CGScheme_GPU.F90:280-296
module MD_CGScheme_GPU
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use mdb_c_binding
  implicit none
contains
  subroutine Do_CG_Forsteps_DEV(SimBox, CtrlParam, ForceClass, MXNUMSTEPS, METH)
      use MD_Forceclass_Register_GPU
      type(SimMDBox), dimension(:)      :: SimBox
      type(SimMDCtrl),       intent(in) :: CtrlParam
      type(MDForceClassGPU), intent(in) :: ForceClass
      integer,               intent(in) :: MXNUMSTEPS, METH
      integer(c_int) :: IFLAG
      real(c_double) :: DELEPOT
      if(mdb_cg(m_CTX, MXNUMSTEPS, METH, CtrlParam%STEEPEST_MxStep*SimBox(1)%RR, CtrlParam%STEEPEST_MiStep*SimBox(1)%RR,   &
                CtrlParam%STEEPEST_MiDelE*CP_EV2ERG, IFLAG, DELEPOT) .lt. 0) stop "MDPSCU Error: mdb_cg failed"
      if(IFLAG .eq. 0) then
         write(*,fmt="(A, I8)")      " MDPSCU WARNING: CG finished after max steps: ", MXNUMSTEPS
         call ONWARNING(gm_OnWarning)
      else
         write(*,fmt="(A, I8)")      " MDPSCU Message: CG finished after steps: ", IFLAG
         write(*,fmt="(A, 1PE13.4)") "                 with max energy uncertainty(ev): ", DELEPOT*CP_ERG2EV
      end if
  end subroutine
end module MD_CGScheme_GPU

!--- CommonGPU/MD_LBFGSScheme_GPU.F90:177-388
module MD_LBFGSScheme_GPU
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MD_Forceclass_Register_GPU
  use mdb_c_binding
  implicit none
contains
  subroutine DO_LBFGSB_FORSTEPS_DEV(SimBox, CtrlParam, ForceClass, MXNUMSTEPS, IFLAG)
      type(SimMDBox), dimension(:)      :: SimBox
      type(SimMDCtrl),       intent(in) :: CtrlParam
      type(MDForceClassGPU), intent(in) :: ForceClass
      integer,               intent(in) :: MXNUMSTEPS
      integer                           :: IFLAG
      integer(c_int) :: NFG, NITER
      if(mdb_lbfgs(m_CTX, MXNUMSTEPS, CtrlParam%LBFGS_MSave, CtrlParam%LBFGS_Factr, CtrlParam%LBFGS_PGtol, IFLAG, NFG, NITER) .lt. 0) &
         stop "MDPSCU Error: mdb_lbfgs failed"
      if(IFLAG .gt. 0) then
         write(*,fmt="(A, I2, A, I7, A)") " MDPSCU Warning: LBFG exit with code ", IFLAG, " after ", MXNUMSTEPS, " steps"
         call ONWARNING(gm_OnWarning)
      end if
  end subroutine
end module MD_LBFGSScheme_GPU
