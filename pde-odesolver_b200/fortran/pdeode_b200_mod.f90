!==============================================================================
!   pdeode_b200 -- ISO_C_BINDING interface to libpdeode_b200.so
!------------------------------------------------------------------------------
!   NEVER COMPILED BY A FORTRAN COMPILER: the image this repository is built in
!   has none (gfortran / nvfortran / flang are all absent).  What is checked
!   instead (tests/test_fortran_binding.py): the C prototypes and the struct
!   layout these interface bodies declare under the standard's interoperability
!   rules are compared with include/pdeode_b200.h by gcc, and the call sites of
!   INTEGRATION.md against these interfaces.  It is the binding a maintainer of
!   the reference adds to keep the Fortran host and run the time step on the
!   B200; INTEGRATION.md shows where solveOrder2 calls it.
!
!   The reference is built with -fdefault-real-8, so its `real` is real(c_double)
!   and its default integer is integer(c_int); arrays are passed as they are.
!==============================================================================
module pdeode_b200
    use iso_c_binding
    implicit none

    integer(c_int), parameter :: PDEODE_OK = 0
    integer(c_int), parameter :: PDEODE_ERR_PICARD_CAP = 4
    integer(c_int), parameter :: PDEODE_ERR_MASK = 255
    integer(c_int), parameter :: PDEODE_WARN_CG_CAP = 256

    ! mirrors pdeode_params (include/pdeode_b200.h)
    type, bind(C) :: pdeode_params
        real(c_double) :: delta, nu, kappa, gama
        real(c_double) :: tDel, xDel
        real(c_double) :: e1, e2, eSoln
        integer(c_int) :: alpha, beta
        integer(c_int) :: dSelect, fSelect, gSelect
        integer(c_int) :: reserved = 0
    end type pdeode_params

    interface
        integer(c_int) function pdeode_create(ctx, row, col, p, device) bind(C, name="pdeode_create")
            import :: c_ptr, c_int, pdeode_params
            type(c_ptr), intent(out) :: ctx
            integer(c_int), value :: row, col, device
            type(pdeode_params), intent(in) :: p
        end function

        ! the same grid as row slabs over ngpus devices of this one process; devices(ngpus)
        integer(c_int) function pdeode_create_group(ctx, row, col, p, ngpus, devices) &
                bind(C, name="pdeode_create_group")
            import :: c_ptr, c_int, pdeode_params
            type(c_ptr), intent(out) :: ctx
            integer(c_int), value :: row, col, ngpus
            type(pdeode_params), intent(in) :: p
            integer(c_int), dimension(*), intent(in) :: devices
        end function

        integer(c_int) function pdeode_destroy(ctx) bind(C, name="pdeode_destroy")
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
        end function

        integer(c_int) function pdeode_set_state(ctx, M, C) bind(C, name="pdeode_set_state")
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: ctx
            real(c_double), dimension(*), intent(in) :: M, C
        end function

        integer(c_int) function pdeode_get_state(ctx, M, C) bind(C, name="pdeode_get_state")
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: ctx
            real(c_double), dimension(*), intent(out) :: M, C
        end function

        ! setInitialConditions on the device (P:262-439); px, py: the npts inoculation points
        integer(c_int) function pdeode_set_initial_condition(ctx, MinitialCond, depth, height, npts, px, py) &
                bind(C, name="pdeode_set_initial_condition")
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: ctx
            integer(c_int), value :: MinitialCond, npts
            real(c_double), value :: depth, height
            real(c_double), dimension(*), intent(in) :: px, py
        end function

        ! frames without moving whole fields: every filter-th row and column (P:829-842)
        integer(c_int) function pdeode_frame_size(ctx, filter, nrows_out, ncols_out) bind(C, name="pdeode_frame_size")
            import :: c_ptr, c_int
            type(c_ptr), value :: ctx
            integer(c_int), value :: filter
            integer(c_int), intent(out) :: nrows_out, ncols_out
        end function

        integer(c_int) function pdeode_get_frame(ctx, filter, M, C) bind(C, name="pdeode_get_frame")
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: ctx
            integer(c_int), value :: filter
            real(c_double), dimension(*), intent(out) :: M, C
        end function

        integer(c_int) function pdeode_get_frame_rowmeans(ctx, filter, meanM, meanC) &
                bind(C, name="pdeode_get_frame_rowmeans")
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: ctx
            integer(c_int), value :: filter
            real(c_double), dimension(*), intent(out) :: meanM, meanC
        end function

        integer(c_int) function pdeode_set_maxit(ctx, maxit) bind(C, name="pdeode_set_maxit")
            import :: c_ptr, c_int, c_long_long
            type(c_ptr), value :: ctx
            integer(c_long_long), value :: maxit
        end function

        integer(c_int) function pdeode_step(ctx, countIters, nitSum, nitMax) bind(C, name="pdeode_step")
            import :: c_ptr, c_int, c_long_long
            type(c_ptr), value :: ctx
            integer(c_int), intent(out) :: countIters
            integer(c_long_long), intent(out) :: nitSum, nitMax
        end function
        integer(c_int) function pdeode_step_many(ctx, n, stepsDone, countSum, countMax, nitSum, nitMax) &
                bind(C, name="pdeode_step_many")
            import :: c_ptr, c_int, c_long_long
            type(c_ptr), value :: ctx
            integer(c_int), value :: n
            integer(c_int), intent(out) :: stepsDone, countMax
            integer(c_long_long), intent(out) :: countSum, nitSum, nitMax
        end function

        integer(c_int) function pdeode_mass(ctx, meanM, meanC) bind(C, name="pdeode_mass")
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(out) :: meanM, meanC
        end function

        integer(c_int) function pdeode_peak(ctx, peak, height, intfac) bind(C, name="pdeode_peak")
            import :: c_ptr, c_int, c_double
            type(c_ptr), value :: ctx
            real(c_double), intent(inout) :: peak, height, intfac
        end function
    end interface
end module pdeode_b200
