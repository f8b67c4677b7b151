// Dormand-Prince data shared by every translation unit that holds stage kernels: liboccm_b200.so (occm_solver.cu) and the
// mechanism-specialised kernel images (spec/occm_spec.cu, compiled per mechanism).
#pragma once

// Dormand-Prince coefficients exactly as scipy builds them (scipy/integrate/_ivp/rk.py:280-320, Python float
// divisions == IEEE double divisions evaluated at compile time here).
namespace dp {
constexpr double C2 = 1.0 / 5, C3 = 3.0 / 10, C4 = 4.0 / 5, C5 = 8.0 / 9;
constexpr double A21 = 1.0 / 5;
constexpr double A31 = 3.0 / 40, A32 = 9.0 / 40;
constexpr double A41 = 44.0 / 45, A42 = -56.0 / 15, A43 = 32.0 / 9;
constexpr double A51 = 19372.0 / 6561, A52 = -25360.0 / 2187, A53 = 64448.0 / 6561, A54 = -212.0 / 729;
constexpr double A61 = 9017.0 / 3168, A62 = -355.0 / 33, A63 = 46732.0 / 5247, A64 = 49.0 / 176, A65 = -5103.0 / 18656;
constexpr double B1 = 35.0 / 384, B3 = 500.0 / 1113, B4 = 125.0 / 192, B5 = -2187.0 / 6784, B6 = 11.0 / 84;
constexpr double E1 = -71.0 / 57600, E3 = 71.0 / 16695, E4 = -71.0 / 1920, E5 = 17253.0 / 339200, E6 = -22.0 / 525, E7 = 1.0 / 40;
// dense output polynomial, rows = stages 1..7 (row 2 is zero), columns = powers x^1..x^4 (rk.py:321-335)
constexpr double P11 = 1.0, P12 = -8048581381.0 / 2820520608, P13 = 8663915743.0 / 2820520608, P14 = -12715105075.0 / 11282082432;
constexpr double P32 = 131558114200.0 / 32700410799, P33 = -68118460800.0 / 10900136933, P34 = 87487479700.0 / 32700410799;
constexpr double P42 = -1754552775.0 / 470086768, P43 = 14199869525.0 / 1410260304, P44 = -10690763975.0 / 1880347072;
constexpr double P52 = 127303824393.0 / 49829197408, P53 = -318862633887.0 / 49829197408, P54 = 701980252875.0 / 199316789632;
constexpr double P62 = -282668133.0 / 205662961, P63 = 2019193451.0 / 616988883, P64 = -1453857185.0 / 822651844;
constexpr double P72 = 40617522.0 / 29380423, P73 = -110615467.0 / 29380423, P74 = 69997945.0 / 29380423;
constexpr double SAFETY = 0.9, MIN_FACTOR = 0.2, MAX_FACTOR = 10.0, ERROR_EXPONENT = -1.0 / 5;
}  // namespace dp

struct OccmRkDev {
    double* Y[2];     // state ping-pong; Y[parity[m]] is the current state of member m
    double* KF[2];    // K1 / K7 ping-pong (first-same-as-last)
    double* K2; double* K3; double* K4; double* K5; double* K6;
    double* YS[2];    // materialised stage states
    double* t; double* h; double* h_abs; double* t_new; double* t_old;
    int* parity; int* status; int* rejected; int* accepted_now; int* out_lo; int* out_hi; int* n_rows;
    long long* n_acc; long long* n_rej; long long* nfev;
    double* err_partial;  // [M][chunks_per_member]
    int* n_running;
    double rtol, atol, t_bound;
    long long max_steps;
};
