// host_check.cpp -- TEST INFRASTRUCTURE.  Compiles the product's per-interface / per-cell device functions
// (hydro-cl-lua_b200/csrc/hb_*.cuh, which are __host__ __device__) with g++ -ffp-contract=off so that the CPU test
// suite can compare them with the oracle without a GPU.  Nothing in the product links or loads this library.
#include "../../hydro-cl-lua_b200/csrc/hb_roe.cuh"
#include "../../hydro-cl-lua_b200/csrc/hb_eqn_euler.cuh"
#include "../../hydro-cl-lua_b200/csrc/hb_eqn_mhd.cuh"
#include "../../hydro-cl-lua_b200/csrc/hb_roe_fast.cuh"

using namespace hb;

template<class Eqn> static void roe(int side, const double* params, const double* UL_, const double* UR_, double* F_) {
	typedef typename Eqn::real real;
	typename Eqn::Params p = Eqn::makeParams(params);
	real UL[Eqn::nI], UR[Eqn::nI], F[Eqn::nI];
	for (int q = 0; q < Eqn::nI; ++q) { UL[q] = real(UL_[q]); UR[q] = real(UR_[q]); }
	if (side == 0) roeFlux<Eqn, 0>(F, p, UL, UR);
	else if (side == 1) roeFlux<Eqn, 1>(F, p, UL, UR);
	else roeFlux<Eqn, 2>(F, p, UL, UR);
	for (int q = 0; q < Eqn::nI; ++q) F_[q] = double(F[q]);
}
template<class Eqn> static void roeLim(int side, const double* params, int lim, double dt_dx, const double* U4, double* F_) {
	typedef typename Eqn::real real;
	typename Eqn::Params p = Eqn::makeParams(params);
	real U[4][Eqn::nI], F[Eqn::nI];
	for (int c = 0; c < 4; ++c) for (int q = 0; q < Eqn::nI; ++q) U[c][q] = real(U4[c * Eqn::nI + q]);
	if (side == 0) roeFluxLimited<Eqn, 0>(F, p, lim, real(dt_dx), U[0], U[1], U[2], U[3]);
	else if (side == 1) roeFluxLimited<Eqn, 1>(F, p, lim, real(dt_dx), U[0], U[1], U[2], U[3]);
	else roeFluxLimited<Eqn, 2>(F, p, lim, real(dt_dx), U[0], U[1], U[2], U[3]);
	for (int q = 0; q < Eqn::nI; ++q) F_[q] = double(F[q]);
}
template<class Eqn> static void constrain(const double* params, double* U_) {
	typedef typename Eqn::real real;
	typename Eqn::Params p = Eqn::makeParams(params);
	real U[Eqn::nI];
	for (int q = 0; q < Eqn::nI; ++q) U[q] = real(U_[q]);
	Eqn::constrainU(p, U);
	for (int q = 0; q < Eqn::nI; ++q) U_[q] = double(U[q]);
}
template<class Eqn> static double dtCell(const double* params, const double* U_, const double* dx_, int dim) {
	typedef typename Eqn::real real;
	typename Eqn::Params p = Eqn::makeParams(params);
	real U[Eqn::nI]; real dx[3] = {real(dx_[0]), real(dx_[1]), real(dx_[2])};
	for (int q = 0; q < Eqn::nI; ++q) U[q] = real(U_[q]);
	return double(Eqn::calcDTCell(p, U, dx, dim));
}

// the solver's flux plug-in as the tile kernel selects it (hb_roe.cuh interfaceFlux: 0 roe, 1 hll, 2 rusanov, 3 euler-hllc)
template<class Eqn> static void ifaceFlux(int side, int flux, int fluxParam, const double* params, const double* UL_, const double* UR_, double* F_) {
	typedef typename Eqn::real real;
	typename Eqn::Params p = Eqn::makeParams(params);
	real UL[Eqn::nI], UR[Eqn::nI], F[Eqn::nI];
	for (int q = 0; q < Eqn::nI; ++q) { UL[q] = real(UL_[q]); UR[q] = real(UR_[q]); }
	if (side == 0) interfaceFlux<Eqn, 0>(flux, fluxParam, F, p, UL, UR);
	else if (side == 1) interfaceFlux<Eqn, 1>(flux, fluxParam, F, p, UL, UR);
	else interfaceFlux<Eqn, 2>(flux, fluxParam, F, p, UL, UR);
	for (int q = 0; q < Eqn::nI; ++q) F_[q] = double(F[q]);
}
// the two-face-state reconstructions of hb_roe.cuh (sp.plmMode as in hb_fv_kernels.cuh)
template<class Eqn, int SIDE> static void plmFacesSide(int mode, int lim, typename Eqn::real dt_dx, typename Eqn::Params const& p,
	typename Eqn::real const (&UL)[Eqn::nI], typename Eqn::real const (&U)[Eqn::nI], typename Eqn::real const (&UR)[Eqn::nI],
	typename Eqn::real (&L)[Eqn::nI], typename Eqn::real (&R)[Eqn::nI]) {
	if (mode == 4) plmPrimFaces<Eqn>(L, R, p, lim, UL, U, UR);
	else if (mode == 5) plmConsFluxFaces<Eqn, SIDE>(L, R, p, lim, dt_dx, UL, U, UR);
	else if (mode == 6) plmEigFaces<Eqn, SIDE>(L, R, p, lim, dt_dx, UL, U, UR);
	else if (mode >= 7 && mode <= 10) plmEigPrimFaces<Eqn, SIDE>(L, R, p, mode == 8 || mode == 10, mode >= 9 ? 1 : 0, dt_dx, UL, U, UR);
	else plmAthenaFaces<Eqn, SIDE>(L, R, p, UL, U, UR, mode == 3 ? 1 : 0);
}
template<class Eqn> static void plmFaces(int side, int mode, int lim, double dt_dx, const double* params, const double* UL_, const double* U_, const double* UR_, double* L_, double* R_) {
	typedef typename Eqn::real real;
	typename Eqn::Params p = Eqn::makeParams(params);
	real UL[Eqn::nI], U[Eqn::nI], UR[Eqn::nI], L[Eqn::nI], R[Eqn::nI];
	for (int q = 0; q < Eqn::nI; ++q) { UL[q] = real(UL_[q]); U[q] = real(U_[q]); UR[q] = real(UR_[q]); }
	if (side == 0) plmFacesSide<Eqn, 0>(mode, lim, real(dt_dx), p, UL, U, UR, L, R);
	else if (side == 1) plmFacesSide<Eqn, 1>(mode, lim, real(dt_dx), p, UL, U, UR, L, R);
	else plmFacesSide<Eqn, 2>(mode, lim, real(dt_dx), p, UL, U, UR, L, R);
	for (int q = 0; q < Eqn::nI; ++q) { L_[q] = double(L[q]); R_[q] = double(R[q]); }
}

#define DISPATCH(call) \
	if (eqn == 0 && rb == 8) { typedef Euler<double> E; call; } \
	else if (eqn == 0) { typedef Euler<float> E; call; } \
	else if (rb == 8) { typedef MHD<double> E; call; } \
	else { typedef MHD<float> E; call; }

extern "C" {
void hc_interface_flux(int eqn, int rb, int side, int flux, int fluxParam, const double* params, const double* UL, const double* UR, double* F) { DISPATCH(ifaceFlux<E>(side, flux, fluxParam, params, UL, UR, F)) }
void hc_plm_faces(int eqn, int rb, int side, int mode, int lim, double dt_dx, const double* params, const double* UL, const double* U, const double* UR, double* L, double* R) { DISPATCH(plmFaces<E>(side, mode, lim, dt_dx, params, UL, U, UR, L, R)) }
void hc_roe_flux(int eqn, int rb, int side, const double* params, const double* UL, const double* UR, double* F) { DISPATCH(roe<E>(side, params, UL, UR, F)) }
void hc_roe_flux_limited(int eqn, int rb, int side, const double* params, int lim, double dt_dx, const double* U4, double* F) { DISPATCH(roeLim<E>(side, params, lim, dt_dx, U4, F)) }
void hc_constrainU(int eqn, int rb, const double* params, double* U) { DISPATCH(constrain<E>(params, U)) }
double hc_calc_dt_cell(int eqn, int rb, const double* params, const double* U, const double* dx, int dim) { double r = 0; DISPATCH(r = dtCell<E>(params, U, dx, dim)) return r; }
double hc_plm_half_slope(int rb, int lim, double UL, double U, double UR) {
	return rb == 8 ? plmHalfSlope<double>(lim, UL, U, UR) : double(plmHalfSlope<float>(lim, float(UL), float(U), float(UR)));
}
double hc_limiter(int id, double r) { return limiter<double>(id, r); }
// production ("fast") forms, hb_roe_fast.cuh (double only; compiled without contraction here)
void hc_euler_roe_flux_fast(int side, const double* params, const double* UL_, const double* UR_, double* F_) {
	typedef Euler<double, true> E;
	E::Params p = E::makeParams(params);
	double UL[5], UR[5], F[5];
	for (int q = 0; q < 5; ++q) { UL[q] = UL_[q]; UR[q] = UR_[q]; }
	if (side == 0) roeFluxAuto<E, 0>(F, p, UL, UR);
	else if (side == 1) roeFluxAuto<E, 1>(F, p, UL, UR);
	else roeFluxAuto<E, 2>(F, p, UL, UR);
	for (int q = 0; q < 5; ++q) F_[q] = F[q];
}
void hc_mhd_roe_flux_fast(int side, const double* params, const double* UL_, const double* UR_, double* F_) {
	typedef MHD<double, true> E;
	E::Params p = E::makeParams(params);
	double UL[8], UR[8], F[8];
	for (int q = 0; q < 8; ++q) { UL[q] = UL_[q]; UR[q] = UR_[q]; }
	if (side == 0) roeFluxAuto<E, 0>(F, p, UL, UR);
	else if (side == 1) roeFluxAuto<E, 1>(F, p, UL, UR);
	else roeFluxAuto<E, 2>(F, p, UL, UR);
	for (int q = 0; q < 8; ++q) F_[q] = F[q];
}
double hc_mhd_finish_cell_fast(const double* params, double* U_, const double* dx_, int dim) {
	typedef MHD<double, true> E;
	E::Params p = E::makeParams(params);
	double U[8]; double dx[3] = {dx_[0], dx_[1], dx_[2]}, invdx[3] = {1. / dx_[0], 1. / dx_[1], 1. / dx_[2]};
	for (int q = 0; q < 8; ++q) U[q] = U_[q];
	double dtCell = HUGE_VAL, rate = 0;
	finishCellAuto<E>(p, U, dx, invdx, dim, true, dtCell, rate);
	for (int q = 0; q < 8; ++q) U_[q] = U[q];
	return rate > 0 ? (1. / rate < dtCell ? 1. / rate : dtCell) : dtCell;
}
double hc_plm_half_slope_fast(int lim, double UL, double U, double UR) {
	return lim == 8 ? plmHalfSlopeT<double, 8, true>(lim, UL, U, UR) : plmHalfSlopeT<double, 18, true>(lim, UL, U, UR);
}
// returns the CFL dt of the cell; U is constrained in place
double hc_euler_finish_cell_fast(const double* params, double* U_, const double* dx_, int dim) {
	typedef Euler<double, true> E;
	E::Params p = E::makeParams(params);
	double U[5]; double dx[3] = {dx_[0], dx_[1], dx_[2]}, invdx[3] = {1. / dx_[0], 1. / dx_[1], 1. / dx_[2]};
	for (int q = 0; q < 5; ++q) U[q] = U_[q];
	double dtCell = HUGE_VAL, rate = 0;
	finishCellAuto<E>(p, U, dx, invdx, dim, true, dtCell, rate);
	for (int q = 0; q < 5; ++q) U_[q] = U[q];
	return rate > 0 ? 1. / rate : dtCell;
}
}
