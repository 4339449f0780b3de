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
      if(CtrlParam%IHDUP .lt. 0) then                     ! variable time step, scheme II (:633-655): the halving loop in one call
         if(MOD(ITIME-CtrlParam%IT0+1, IABS(CtrlParam%IHDUP)) .eq. 0) then
            if(mdb_timestep_limit(m_CTX, CtrlParam%HMX, CtrlParam%DMX, CtrlParam%H) .lt. 0) stop "MDPSCU Error: mdb_timestep_limit failed"
         end if
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

module MD_FS_Force_Table_GPU             ! replaces MDLIB/sor/CommonGPU/MD_FS_ForceTable_GPU.F90 (the FS_TYPE slots of Register_ForceClass)
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MD_TYPEDEF_ForceTable
  use MDB_C_BINDING
  implicit none
contains
  subroutine INITIALIZE_FS_Force_Table_DEV(SimBox, CtrlParam, FTable, RelaseTable, MULTIBOX)    ! :178-310
    type(SimMDBox),     intent(inout)::SimBox
    type(SimMDCtrl),    intent(in)   ::CtrlParam
    type(MDForceTable), intent(in)   ::FTable
    integer,            optional     ::RelaseTable, MULTIBOX
      ! FS_TYPE: same tables, embedding -sqrt(rho) evaluated in the kernels (FEMBD / DFEMBD are not read)
      if(mdb_tables_set(m_CTX, MDB_POT_FS, size(FTable%POTR,1), size(FTable%POTR,2), FTable%CSI,                &
                        FTable%POTR, FTable%FPOTR, FTable%POTB, FTable%FPOTB,                                    &
                        size(FTable%FEMBD,1), size(FTable%FEMBD,2), FTable%RHOD, FTable%FEMBD, FTable%DFEMBD,    &
                        FTable%KPAIR, FTable%KEMBD, maxval(CtrlParam%RU*CtrlParam%RU)) .lt. 0)                   &
         stop "MDPSCU Error: mdb_tables_set failed"
  end subroutine
  subroutine CALFORCE_FS_Force_Table2A_DEV(SimBox, CtrlParam)                                    ! :934-1000
    type(SimMDBox),  intent(inout)::SimBox
    type(SimMDCtrl), intent(in)   ::CtrlParam
    real(c_double)::VT(9)
      if(mdb_force(m_CTX, MDB_FORCE, VT) .lt. 0) stop "MDPSCU Error: mdb_force failed"
  end subroutine
  subroutine CALPTENSOR_FS_Force_Table2A_DEV(SimBox, CtrlParam)                                  ! :1348-1412
    type(SimMDBox),  intent(inout)::SimBox
    type(SimMDCtrl), intent(in)   ::CtrlParam
    real(c_double)::VT(9)
      if(mdb_force(m_CTX, ior(MDB_FORCE, MDB_VIRIAL), VT) .lt. 0) stop "MDPSCU Error: mdb_force failed"
      SimBox%VTENSOR = reshape(VT, (/3,3/))
  end subroutine
  subroutine UpdateEPOT_FS_Force_Table2A_DEV(SimBox, CtrlParam)                                  ! :1679-1716
    type(SimMDBox),  intent(inout)::SimBox
    type(SimMDCtrl), intent(in)   ::CtrlParam
    real(c_double)::VT(9)
      if(mdb_force(m_CTX, MDB_EPOT, VT) .lt. 0) stop "MDPSCU Error: mdb_force failed"
  end subroutine
  subroutine CALEPOT_FS_Force_Table2A_DEV(SimBox, CtrlParam)                                     ! :1720-1760: + copy to hm_EPOT
    use MD_Globle_Variables_GPU, only: hm_EPOT
    type(SimMDBox),  intent(inout)::SimBox
    type(SimMDCtrl), intent(in)   ::CtrlParam
    real(c_double)::VT(9)
      if(mdb_force(m_CTX, MDB_EPOT, VT) .lt. 0) stop "MDPSCU Error: mdb_force failed"
      if(mdb_state_download(m_CTX, MDB_F_EPOT, c_loc(hm_EPOT), MDB_ORDER_CELL) .lt. 0) stop "MDPSCU Error: mdb_state_download failed"
  end subroutine
  subroutine CALDEN_FS_Force_Table2A_DEV(SimBox, CtrlParam)                                      ! :875-930
    type(SimMDBox),  intent(inout)::SimBox
    type(SimMDCtrl), intent(in)   ::CtrlParam
    real(c_double)::VT(9)
      if(mdb_force(m_CTX, MDB_DEN, VT) .lt. 0) stop "MDPSCU Error: mdb_force failed"
  end subroutine
  subroutine Cal_FS_AtomicStressTensor_DEV(IDEV, dAVP)                                           ! :2044-2061 (pCalAVStress)
    integer, intent(in)::IDEV
    real(KINDDF), device, dimension(:,:), intent(out)::dAVP
      if(mdb_atomic_stress(m_CTX, c_devloc(dAVP)) .lt. 0) stop "MDPSCU Error: mdb_atomic_stress failed"
  end subroutine
  subroutine Clear_FS_Force_Table_DEV()                                                          ! :314-337
    integer(c_int)::ERR
      ERR = mdb_tables_clear(m_CTX)
  end subroutine
end module MD_FS_Force_Table_GPU

module MSM_MultiGPU_Basic_B200           ! the device bookkeeping of MSMLIB/sor/CommonGPU/MSM_MultiGPU_Basic.F90 the hot path uses
  use MDB_C_BINDING
  implicit none
  integer, parameter :: m_MXDEVICE = 8   ! one box of 8 B200 (the reference allows 6, :22); ONE device per host process here
  integer            :: m_NDEVICE = 0, m_DEVICES(m_MXDEVICE) = -1
contains
  subroutine Initialize_DEVICES(FIRSTDEV, NDEV)                                                  ! :571-599
    integer::FIRSTDEV, NDEV
      if(NDEV .gt. m_MXDEVICE) then
         write(*,*) "MDPSCU Error: the number of devices larger than permitted value ", m_MXDEVICE
         stop
      end if
      if(NDEV .gt. 1) then
         ! several devices of one process is the reference's scheme; here a process drives ONE device and several
         ! processes share a box through mdb_dd_* (slab decomposition over NCCL) or own independent boxes
         write(*,*) "MDPSCU Warning: libmdpscu_b200 drives one device per process; using device ", FIRSTDEV
      end if
      if(mdb_device_count() .le. FIRSTDEV) stop "MDPSCU Error: no such CUDA device"
      if(c_associated(m_CTX)) call mdb_ctx_destroy(m_CTX)
      if(mdb_ctx_create(FIRSTDEV, m_CTX) .lt. 0) stop "MDPSCU Error: mdb_ctx_create failed"
      m_NDEVICE = 1
      m_DEVICES(1) = FIRSTDEV
  end subroutine
  subroutine End_DEVICES()                                                                       ! :640-657
      if(c_associated(m_CTX)) call mdb_ctx_destroy(m_CTX)
      m_CTX = c_null_ptr
      m_NDEVICE = 0
  end subroutine
end module MSM_MultiGPU_Basic_B200

module MD_Globle_Variables_GPU           ! replaces MDLIB/sor/CommonGPU/MD_Globle_Variables_GPU.F90 (:196-510 public names)
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MDB_C_BINDING
  implicit none
  ! host mirrors the application shell reads (reference :113-139); filled by the Copy...From_Devices_to_Host family
  integer                                       :: m_NAPDEV = 0, dm_NPRT = 0
  integer,      dimension(:),   allocatable, target :: hm_ITYP, hm_STATU, hm_GID, hm_GIDINV
  real(KINDDF), dimension(:,:), allocatable, target :: hm_XP, hm_XP1, hm_FP, hm_DIS
  real(KINDDF), dimension(:),   allocatable, target :: hm_EPOT, hm_EKIN
  interface Initialize_Globle_Variables_DEV
     module procedure Initialize_GB_A_DEV
     module procedure Initialize_GB_B_DEV
  end interface
contains
  subroutine Initialize_GB_A_DEV(SimBox, CtrlParam)                                              ! :585-655: SimBox(:) = MULTIBOX
    type(SimMDBox), dimension(:)::SimBox
    type(SimMDCtrl)             ::CtrlParam
    integer::NB, NPRT, I, IS
      NB = size(SimBox); NPRT = SimBox(1)%NPRT
      call Clear_Globle_Variables_DEV()
      if(mdb_box_set(m_CTX, NB, NPRT, SimBox(1)%BOXLOW, SimBox(1)%ZL, reshape(SimBox(1)%BOXSHAPE, (/9/)), CtrlParam%IFPD, &
                     SimBox(1)%NGROUP, SimBox(1)%CM) .lt. 0) stop "MDPSCU Error: mdb_box_set failed"
      dm_NPRT = NB*NPRT; m_NAPDEV = dm_NPRT
      allocate(hm_ITYP(dm_NPRT), hm_STATU(dm_NPRT), hm_GID(dm_NPRT), hm_GIDINV(dm_NPRT), hm_XP(dm_NPRT,3), hm_XP1(dm_NPRT,3), &
               hm_FP(dm_NPRT,3), hm_DIS(dm_NPRT,3), hm_EPOT(dm_NPRT), hm_EKIN(dm_NPRT))
      IS = 0
      do I=1, NB                                                                                 ! :628-640
         hm_ITYP(IS+1:IS+NPRT) = SimBox(I)%ITYP;      hm_STATU(IS+1:IS+NPRT) = SimBox(I)%STATU
         hm_XP(IS+1:IS+NPRT,:) = SimBox(I)%XP;        hm_XP1(IS+1:IS+NPRT,:) = SimBox(I)%XP1
         hm_DIS(IS+1:IS+NPRT,:)= SimBox(I)%DIS;       hm_FP(IS+1:IS+NPRT,:)  = SimBox(I)%FP
         IS = IS + NPRT
      end do
      call CopyAllFrom_Host_to_Devices()
  end subroutine
  subroutine Initialize_GB_B_DEV(SimBox, CtrlParam)                                              ! one box
    type(SimMDBox) ::SimBox
    type(SimMDCtrl)::CtrlParam
    type(SimMDBox), dimension(1)::SB
      SB(1) = SimBox
      call Initialize_GB_A_DEV(SB, CtrlParam)
  end subroutine
  subroutine Clear_Globle_Variables_DEV()                                                        ! :658-714
      if(allocated(hm_ITYP)) deallocate(hm_ITYP, hm_STATU, hm_GID, hm_GIDINV, hm_XP, hm_XP1, hm_FP, hm_DIS, hm_EPOT, hm_EKIN)
      dm_NPRT = 0; m_NAPDEV = 0
  end subroutine
  subroutine CopyAllFrom_Host_to_Devices()                                                       ! :1076-1122
      if(mdb_state_upload(m_CTX, MDB_F_XP,    c_loc(hm_XP),    MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: upload XP"
      if(mdb_state_upload(m_CTX, MDB_F_XP1,   c_loc(hm_XP1),   MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: upload XP1"
      if(mdb_state_upload(m_CTX, MDB_F_DIS,   c_loc(hm_DIS),   MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: upload DIS"
      if(mdb_state_upload(m_CTX, MDB_F_FP,    c_loc(hm_FP),    MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: upload FP"
      if(mdb_state_upload(m_CTX, MDB_F_ITYP,  c_loc(hm_ITYP),  MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: upload ITYP"
      if(mdb_state_upload(m_CTX, MDB_F_STATU, c_loc(hm_STATU), MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: upload STATU"
  end subroutine
  subroutine CopyAllFrom_Devices_to_Host()                                                       ! :1023-1072 (original order restored)
      call CopyXPFrom_Devices_to_Host();   call CopyXP1From_Devices_to_Host();  call CopyFPFrom_Devices_to_Host()
      call CopyDISFrom_Devices_to_Host();  call CopyEPOTFrom_Devices_to_Host(); call CopyEKINFrom_Devices_to_Host()
      call CopyStatuFrom_Devices_to_Host()
  end subroutine
  subroutine CopyXPFrom_Devices_to_Host()
      if(mdb_state_download(m_CTX, MDB_F_XP, c_loc(hm_XP), MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: download XP"
  end subroutine
  subroutine CopyXP1From_Devices_to_Host()
      if(mdb_state_download(m_CTX, MDB_F_XP1, c_loc(hm_XP1), MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: download XP1"
  end subroutine
  subroutine CopyFPFrom_Devices_to_Host()
      if(mdb_state_download(m_CTX, MDB_F_FP, c_loc(hm_FP), MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: download FP"
  end subroutine
  subroutine CopyDISFrom_Devices_to_Host()
      if(mdb_state_download(m_CTX, MDB_F_DIS, c_loc(hm_DIS), MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: download DIS"
  end subroutine
  subroutine CopyEPOTFrom_Devices_to_Host()
      if(mdb_state_download(m_CTX, MDB_F_EPOT, c_loc(hm_EPOT), MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: download EPOT"
  end subroutine
  subroutine CopyEKINFrom_Devices_to_Host()
      if(mdb_ekin(m_CTX) .lt. 0) stop "MDPSCU Error: mdb_ekin failed"
      if(mdb_state_download(m_CTX, MDB_F_EKIN, c_loc(hm_EKIN), MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: download EKIN"
  end subroutine
  subroutine CopyStatuFrom_Devices_to_Host()
      if(mdb_state_download(m_CTX, MDB_F_STATU, c_loc(hm_STATU), MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: download STATU"
      if(mdb_state_download(m_CTX, MDB_F_GID, c_loc(hm_GID), MDB_ORDER_CELL) .lt. 0) stop "MDPSCU Error: download GID"
      if(mdb_state_download(m_CTX, MDB_F_GIDINV, c_loc(hm_GIDINV), MDB_ORDER_CELL) .lt. 0) stop "MDPSCU Error: download GIDINV"
  end subroutine
  subroutine CopyXPFrom_Devices_to_Host1(hXP)                                                    ! :2133-2147
    real(KINDDF), dimension(:,:), target::hXP
      if(mdb_state_download(m_CTX, MDB_F_XP, c_loc(hXP), MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: download XP"
  end subroutine
  subroutine Synchroniz_XP_on_Devices()                                                          ! :2026-2040
      ! one device per process: nothing to gather (the decomposed run exchanges ghost layers inside mdb_dd_run)
  end subroutine
  subroutine SynchronizeDevices()
      if(mdb_sync(m_CTX) .lt. 0) stop "MDPSCU Error: mdb_sync failed"
  end subroutine
end module MD_Globle_Variables_GPU

module MD_SimBoxArray_GPU                ! replaces MDLIB/sor/CommonGPU/MD_SimBoxArray_GPU.F90 (:24-88 generic names)
  use MD_TYPEDEF_SimMDBox
  use MD_Globle_Variables_GPU
  use MDB_C_BINDING
  implicit none
  interface CopyIn_SimBox_DEV
     module procedure CopyIn_SimBoxA
  end interface
  interface CopyOut_SimBox_DEV
     module procedure CopyOut_SimBoxA
  end interface
contains
  subroutine CopyIn_SimBoxA(SimBox)                                                              ! :91-155
    type(SimMDBox), dimension(:)::SimBox
    integer::I, IS, NPRT
      NPRT = SimBox(1)%NPRT; IS = 0
      do I=1, size(SimBox)
         hm_XP(IS+1:IS+NPRT,:) = SimBox(I)%XP;   hm_XP1(IS+1:IS+NPRT,:) = SimBox(I)%XP1
         hm_DIS(IS+1:IS+NPRT,:)= SimBox(I)%DIS;  hm_FP(IS+1:IS+NPRT,:)  = SimBox(I)%FP
         hm_STATU(IS+1:IS+NPRT)= SimBox(I)%STATU; hm_ITYP(IS+1:IS+NPRT) = SimBox(I)%ITYP
         IS = IS + NPRT
      end do
      call CopyAllFrom_Host_to_Devices()
  end subroutine
  subroutine CopyOut_SimBoxA(SimBox)                                                             ! :202-240
    type(SimMDBox), dimension(:)::SimBox
    integer::I, IS, NPRT
      call CopyAllFrom_Devices_to_Host()
      NPRT = SimBox(1)%NPRT; IS = 0
      do I=1, size(SimBox)
         SimBox(I)%XP  = hm_XP(IS+1:IS+NPRT,:);   SimBox(I)%XP1  = hm_XP1(IS+1:IS+NPRT,:)
         SimBox(I)%DIS = hm_DIS(IS+1:IS+NPRT,:);  SimBox(I)%FP   = hm_FP(IS+1:IS+NPRT,:)
         SimBox(I)%EPOT= hm_EPOT(IS+1:IS+NPRT);   SimBox(I)%EKIN = hm_EKIN(IS+1:IS+NPRT)
         SimBox(I)%STATU = hm_STATU(IS+1:IS+NPRT)
         IS = IS + NPRT
      end do
  end subroutine
end module MD_SimBoxArray_GPU

module MD_NeighborsList_GPU_More         ! the remaining public entry points of MD_NeighborsList_GPU.F90 (:116-202)
  use MD_NeighborsList
  use MDB_C_BINDING
  implicit none
contains
  subroutine Copyout_NeighboreList_A2(List)                                                      ! :584-600: cell-sorted order
    type(NEIGHBOR_LIST), target::List
      if(mdb_nlist_copyout(m_CTX, List%KVOIS, List%INDI, MDB_ORDER_CELL) .lt. 0) stop "MDPSCU Error: mdb_nlist_copyout failed"
  end subroutine
  subroutine Copyout_NeighboreList_ORIG(List)                                                    ! :465-560 with GID: original order
    type(NEIGHBOR_LIST), target::List
      if(mdb_nlist_copyout(m_CTX, List%KVOIS, List%INDI, MDB_ORDER_ORIGINAL) .lt. 0) stop "MDPSCU Error: mdb_nlist_copyout failed"
  end subroutine
  subroutine GetCellInform(TNC, NCELL, MXNAC)                                                    ! :2391-2399
    integer::TNC, NCELL(3), MXNAC
      if(mdb_nlist_cellinfo(m_CTX, NCELL, TNC, MXNAC) .lt. 0) stop "MDPSCU Error: mdb_nlist_cellinfo failed"
  end subroutine
  subroutine Clear_NeighboreList_DEV()                                                           ! :139-152
    integer(c_int)::ERR
      ERR = mdb_nlist_clear(m_CTX)
  end subroutine
end module MD_NeighborsList_GPU_More

module MD_LocalTempMethod_GPU            ! replaces MDLIB/sor/LocalTempCtrlMeths/MD_LocalTempMethod_GPU.F90 (:95-138) for the EPC part
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MDB_C_BINDING
  implicit none
contains
  subroutine Do_ResetParam_DEV(SimBox, CtrlParam)                                                ! -> Reset_EPCMOD_DEV, EPC/MD_EP_Coupling_GPU.F90:370-417
    type(SimMDBox),  intent(in)::SimBox
    type(SimMDCtrl), intent(in)::CtrlParam
    integer(c_int)::ENABLE(MDB_MXGROUP)
    real(c_double)::TE(MDB_MXGROUP), ALPHA(MDB_MXGROUP), CUT(MDB_MXGROUP), HE(MDB_MXGROUP)
    integer::I
      ENABLE = 0; TE = 0; ALPHA = 1; CUT = 0; HE = 0
      do I=1, SimBox%NGROUP                                                                      ! :384-391: V2TI, EPA, EPACUT, EPUPPER are formed inside mdb_epc_set
         ENABLE(I) = iand(CtrlParam%LT_CTRL(I)%METH, CP_TICTRL_METH_EPC)
         TE(I)     = CtrlParam%LT_CTRL(I)%TI
         ALPHA(I)  = CtrlParam%LT_CTRL(I)%EPC_Alpha
         CUT(I)    = CtrlParam%LT_CTRL(I)%EPC_CUT
         HE(I)     = CtrlParam%LT_CTRL(I)%EPC_HE
      end do
      if(mdb_epc_set(m_CTX, ENABLE, TE, ALPHA, CUT, HE) .lt. 0) stop "MDPSCU Error: mdb_epc_set failed"
  end subroutine
  subroutine Do_EPCForce_DEV(SimBox, CtrlParam)                                                  ! :119-138 -> Do_EPCMOD_DEV :421-493
    type(SimMDBox), dimension(:)::SimBox
    type(SimMDCtrl)             ::CtrlParam
      if(mdb_epc_apply(m_CTX) .lt. 0) stop "MDPSCU Error: mdb_epc_apply failed"
  end subroutine
end module MD_LocalTempMethod_GPU

module MD_Method_ParRep_GPU_EventDetect  ! the device pieces of Do_ChangeDetect, Appshell/MD_Method_ParRep_GPU.F90:1094-1167
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MDB_C_BINDING
  implicit none
contains
  ! the body of Do_ChangeDetect between "allocate(SwapBox ...)" and "deallocate": no host copies of the replicas
  subroutine Do_ChangeDetect_DEV(SimBoxIni, NB, CtrlParam, Mask, IBT, NCB, FlagBox)
    type(SimMDBox),  intent(in) ::SimBoxIni
    integer,         intent(in) ::NB
    type(SimMDCtrl), intent(in) ::CtrlParam
    integer, dimension(:), intent(in) ::Mask
    integer, intent(out)::IBT, NCB
    integer, dimension(:)::FlagBox
    integer(c_int)::IFLAG
    real(c_double)::MAXMOVE, DELEPOT
    integer(c_int), target::DUMMY(1)
      if(mdb_state_save(m_CTX) .lt. 0) stop "MDPSCU Error: mdb_state_save failed"
      ! Do_Damp(SwapBox, m_CtrlParamDamp, gm_ForceClass): the "ST" scheme; the other schemes call mdb_cg / mdb_lbfgs / mdb_dyndamp
      if(mdb_steepest(m_CTX, CtrlParam%Quench_Steps, 0, CtrlParam%STEEPEST_Alpha, CtrlParam%STEEPEST_MxStep*SimBoxIni%RR,  &
                      CtrlParam%STEEPEST_MiStep*SimBoxIni%RR, CtrlParam%STEEPEST_MiDelE*CP_EVERG, IFLAG, MAXMOVE, DELEPOT) .lt. 0) &
         stop "MDPSCU Error: mdb_steepest failed"
      if(mdb_compare(m_CTX, SimBoxIni%XP, Mask, CtrlParam%STRCUT_DRTol, FlagBox, DUMMY, IBT, NCB) .lt. 0) &
         stop "MDPSCU Error: mdb_compare failed"
      if(mdb_state_restore(m_CTX) .lt. 0) stop "MDPSCU Error: mdb_state_restore failed"
  end subroutine
end module MD_Method_ParRep_GPU_EventDetect

module MD_ST_Coupling_GPU                ! replaces MDLIB/sor/LocalTempCtrlMeths/Stopping/MD_ST_Coupling_GPU.F90 (:317-427, :1265-1291)
  use MD_CONSTANTS
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MD_TYPEDEF_StopRange_Table
  use MDB_C_BINDING
  implicit none
  type(MDSTPTable), pointer, private::m_pSTPTable => null()
contains
  subroutine Initialize_STMOD_DEV(SimBox, CtrlParam, STPTable)                                   ! :334-357
    type(SimMDBox),  intent(in)::SimBox
    type(SimMDCtrl), intent(in)::CtrlParam
    type(MDSTPTable),intent(in), target::STPTable
      m_pSTPTable => STPTable
      call Reset_STMOD_DEV(SimBox, CtrlParam)
  end subroutine
  subroutine Reset_STMOD_DEV(SimBox, CtrlParam)                                                  ! :361-427
    type(SimMDBox)  ::SimBox
    type(SimMDCtrl) ::CtrlParam
    integer(c_int)::ENABLE(MDB_MXGROUP), LOCAL
    real(c_double)::MDEN(MDB_MXGROUP)
    integer::I
      ENABLE = 0; MDEN = 0; LOCAL = 0
      do I=1, SimBox%NGROUP
         ENABLE(I) = iand(CtrlParam%LT_CTRL(I)%METH, CP_TICTRL_METH_ST)
      end do
      if(CtrlParam%ST_CTRL%MDEN .gt. 0.000001D0) then                                            ! :377-379: the given density
         MDEN(1:SimBox%NGROUP) = CtrlParam%ST_CTRL%MDEN*dble(SimBox%NA(1:SimBox%NGROUP))/dble(SimBox%NPRT)
      else if(dabs(CtrlParam%ST_CTRL%MDEN) .le. 0.00001D0) then                                  ! :381-386: atoms of the box / box volume
         MDEN(1:SimBox%NGROUP) = dble(SimBox%NA(1:SimBox%NGROUP))/(SimBox%ZL(1)*SimBox%ZL(2)*SimBox%ZL(3))
      else                                                                                       ! :395-397: mp_STMOD_L
         LOCAL = 1
      end if
      ! ETAB / STAB / KPAIR exactly as type(MDSTPTable) holds them (E(0:NTAB), STPWR(0:NTAB,NK), KPAIR(NG,NG))
      if(mdb_stopping_set(m_CTX, size(m_pSTPTable%E), size(m_pSTPTable%STPWR, dim=2), m_pSTPTable%E, m_pSTPTable%STPWR, &
                          m_pSTPTable%KPAIR, ENABLE, MDEN) .lt. 0) stop "MDPSCU Error: mdb_stopping_set failed"
      if(mdb_stopping_options(m_CTX, LOCAL, CtrlParam%ST_CTRL%SaveEloss) .lt. 0) stop "MDPSCU Error: mdb_stopping_options failed"
  end subroutine
  subroutine Do_STMOD_DEV(SimBox, CtrlParam)                                                     ! :1265-1291
    type(SimMDBox), dimension(:)::SimBox
    type(SimMDCtrl)             ::CtrlParam
      if(mdb_stopping_apply(m_CTX, CtrlParam%H) .lt. 0) stop "MDPSCU Error: mdb_stopping_apply failed"
  end subroutine
  subroutine CopyElossFrom_Devices_to_Host(hELOSS)                                               ! :1295-1312 (accumulated, original order, erg)
    real(KINDDF), dimension(:)::hELOSS
      if(mdb_stopping_eloss(m_CTX, hELOSS, 0) .lt. 0) stop "MDPSCU Error: mdb_stopping_eloss failed"
  end subroutine
  subroutine Clear_STMOD_DEV(SimBox)                                                             ! :317-330
    type(SimMDBox), optional::SimBox
    integer(c_int)::IDUM(1)
    real(c_double)::DDUM(1)
      if(mdb_stopping_set(m_CTX, 0, 0, DDUM, DDUM, IDUM, IDUM, DDUM) .lt. 0) stop "MDPSCU Error: mdb_stopping_set failed"
  end subroutine
end module MD_ST_Coupling_GPU

module MD_ActiveRegion_GPU               ! replaces MDLIB/sor/CommonGPU/MD_ActiveRegion_GPU.F90 (:193-238, :283-424, :1329-1353)
  use MD_CONSTANTS
  use MD_TYPEDEF_SimMDBox
  use MD_TYPEDEF_SimMDCtrl
  use MDB_C_BINDING
  implicit none
contains
  subroutine Initialize_ActiveRegion_DEV(SimBox, CtrlParam)                                      ! :193-215: nothing to allocate here
    type(SimMDBox), dimension(:), intent(in) ::SimBox
    type(SimMDCtrl),              intent(in) ::CtrlParam
  end subroutine
  subroutine ActivateRegion_DEV(SimBox, CtrlParam)                                               ! :1329-1353
    type(SimMDBox), dimension(:) ::SimBox
    type(SimMDCtrl)              ::CtrlParam
    integer(c_int)::METHOD, NACT
      if(iand(CtrlParam%AR_METHOD, CP_ENABLE_AR) .ne. CP_ENABLE_AR) return
      METHOD = 0
      if(iand(CtrlParam%AR_METHOD, CP_CENTPART_AR) .eq. CP_CENTPART_AR) METHOD = METHOD + 1
      if(iand(CtrlParam%AR_METHOD, CP_EKIN_AR)     .eq. CP_EKIN_AR)     METHOD = METHOD + 2
      if(iand(CtrlParam%AR_METHOD, CP_KEEP_AR)     .eq. CP_KEEP_AR)     METHOD = METHOD + 4
      if(ibits(CtrlParam%AR_METHOD, CP_BYNBSETBIT_AR, 1) .gt. 0)        METHOD = METHOD + 8       ! ActiveByNeigbors1, else ActiveByCells1
      NACT = mdb_active_region(m_CTX, METHOD, CtrlParam%AR_CENTPART, CtrlParam%AR_EKIN*CP_EVERG, CtrlParam%AR_Extend)
      if(NACT .lt. 0) stop "MDPSCU Error: mdb_active_region failed"
  end subroutine
  subroutine Active_All_ActiveRegion_DEV(SimBox, CtrlParam)                                      ! :408-424
    type(SimMDBox), dimension(:) ::SimBox
    type(SimMDCtrl)              ::CtrlParam
      if(mdb_active_all(m_CTX, 1) .lt. 0) stop "MDPSCU Error: mdb_active_all failed"
  end subroutine
  subroutine DeActive_All_ActiveRegion_DEV(SimBox, CtrlParam)                                    ! :314-330
    type(SimMDBox), dimension(:) ::SimBox
    type(SimMDCtrl)              ::CtrlParam
      if(mdb_active_all(m_CTX, 0) .lt. 0) stop "MDPSCU Error: mdb_active_all failed"
  end subroutine
  subroutine Clear_ActiveRegion_DEV()                                                            ! :219-238
  end subroutine
end module MD_ActiveRegion_GPU
