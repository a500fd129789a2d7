--[[
hydrob200/fvsolver.lua -- coarse-grained drop-in: a FiniteVolumeSolver whose update path runs the fused B200 kernels.

Usage in config.lua (instead of `require 'hydro.solver.fvsolver'`):

	self.solvers:insert(require 'hydrob200.fvsolver'(args))

The class derives from the reference's own hydro/solver/fvsolver.lua, so the equation / flux / integrator / init-cond
composition, the GUI vars and the `solver:update()` / `solver:calcDT()` / `solver:step(dt)` / `solver:boundary()` API are
the reference's.  Only the methods that enqueue OpenCL work are overridden (SURVEY.md 8b "coarse-grained"):

	SolverBase:update()      hydro/solver/solverbase.lua:3026-3190  -> hb_fv_update(fv, 1)   (dt stays on the device)
	SolverBase:calcDT()      :3004-3023                             -> hb_fv_calc_dt
	SolverBase:step(dt)      :3193-3238                             -> hb_fv_step
	GridSolver:boundary()    hydro/solver/gridsolver.lua:1316       -> hb_fv_boundary
	SolverBase:constrainU()  :2116-2127                             -> hb_fv_constrainU
	FiniteVolumeSolver:calcDeriv(derivBufObj, dt)  hydro/solver/fvsolver.lua:225-302 -> hb_fv_calc_deriv
	UBufObj:toCPU()/fromCPU()  (display, save, initial condition)   -> hb_fv_get_state / hb_fv_set_state (AoS cons_t)

NOT executed in the authoring container (no LuaJIT); the Python mirror hydro-cl-lua_b200/hydro/solver/fvsolver.py
makes the same calls in the same order and is what the parity tests drive.
--]]
local hb = require 'hydrob200.ffi'
local ffi, lib, check = hb.ffi, hb.lib, hb.check
local FiniteVolumeSolver = require 'hydro.solver.fvsolver'

local B200Solver = FiniteVolumeSolver:subclass()
B200Solver.name = 'fvsolver_b200'

local bcIds = {periodic = lib.HB_BC_PERIODIC, mirror = lib.HB_BC_MIRROR, freeflow = lib.HB_BC_FREEFLOW, none = lib.HB_BC_NONE,
	linear = lib.HB_BC_LINEAR, quadratic = lib.HB_BC_QUADRATIC, fixed = lib.HB_BC_FIXED}
local eqnIds = {euler = lib.HB_EQN_EULER, mhd = lib.HB_EQN_MHD}

function B200Solver:refreshSolverProgram()
	-- no OpenCL program: build the descriptor of the fused path from the solver's own fields
	local d = ffi.new'hb_fv_desc'
	d.eqn = assert(eqnIds[self.eqn.name], "hydrob200: equation not built: "..tostring(self.eqn.name))
	d.dim = self.dim
	for i = 0, 2 do
		d.n[i] = tonumber(self.sizeWithoutBorder.s[i]); d.global_n[i] = d.n[i]
		d.mins[i] = self.mins.s[i]; d.maxs[i] = self.maxs.s[i]
	end
	-- 'plm athena' follows plm.cl:782-879 literally (faces as the tree assigns them, :877-878); use_plm = 3 selects L = left, R = right
	local plmIds = {['piecewise constant'] = 0, ['plm cons'] = 1, ['plm athena'] = 2, ['plm prim'] = 4, ['plm cons with flux'] = 5,
		['plm eig'] = 6, ['plm eig prim'] = 7, ['plm eig prim ref'] = 8}
	d.use_plm = self.usePLM and assert(plmIds[self.usePLM], "hydrob200: usePLM not built: "..tostring(self.usePLM)) or 0
	d.slope_limiter = self.slopeLimiter - 1            -- hydro/app.lua:614-635 is 1-based
	d.flux_limiter = self.fluxLimiter - 1
	-- hydro/flux/*.lua: the flux plug-in object carries its name ('roe', 'hll', 'rusanov')
	d.flux_param = self.flux.hllcMethod or 0          -- hydro/flux/euler-hllc.lua:17
	d.flux = assert(({roe = lib.HB_FLUX_ROE, hll = lib.HB_FLUX_HLL, rusanov = lib.HB_FLUX_RUSANOV, ['euler-hllc'] = lib.HB_FLUX_EULER_HLLC})[self.flux.name],
		"hydrob200: flux not built: "..tostring(self.flux.name))
	if d.flux ~= lib.HB_FLUX_ROE then d.flux_limiter = 0 end   -- only roe usesFluxLimiter (hydro/flux/roe.lua)
	local sides = {'xmin', 'xmax', 'ymin', 'ymax', 'zmin', 'zmax'}
	for i, s in ipairs(sides) do d.bc[i-1] = assert(bcIds[self.boundaryMethods[s].name], s) end
	local int = self.integrator
	d.rk_order = int.order or 0                        -- hydro/int/rk.lua: alphas/betas; hydro/int/fe.lua: order 0
	for i = 1, d.rk_order do for k = 1, d.rk_order do
		d.alphas[(i-1)*d.rk_order + k-1] = int.alphas[i][k] or 0
		d.betas[(i-1)*d.rk_order + k-1] = int.betas[i][k] or 0
	end end
	d.cfl = self.cfl
	d.use_fixed_dt = self.useFixedDT and 1 or 0
	d.fixed_dt = self.fixedDT or 0
	local p = self.eqn.guiVars
	if self.eqn.name == 'euler' then
		d.eqn_params[0], d.eqn_params[1], d.eqn_params[2] = p.heatCapacityRatio.value, p.rhoMin.value, p.PMin.value
	else
		d.eqn_params[0], d.eqn_params[1] = p.heatCapacityRatio.value, p.mu0.value * p.coulomb.value^2   -- mu0 / unit_kg_m_per_C2
	end
	d.use_ctu = self.useCTU and 1 or 0                  -- hydro/solver/gridsolver.lua:102-115
	d.use_graph = 1
	local h = ffi.new'hb_fv*[1]'
	check(lib.hb_fv_create(self.app.env.ctx, d, h), 'hb_fv_create')
	self.fv = ffi.gc(h[0], lib.hb_fv_destroy)
	-- 'fixed' faces (gridsolver.lua:746-764): the Boundary object carries the state its fixedCode writes as .fixedState (numStates reals);
	-- the reference's own uses are cell-independent (init/euler.lua:1859-1879, eqn/einstein.lua:62-80)
	for i, s in ipairs(sides) do
		local b = self.boundaryMethods[s]
		if b.name == 'fixed' then
			local U = ffi.new('double[?]', self.eqn.numStates, assert(b.fixedState, "hydrob200: boundary 'fixed' needs .fixedState"))
			check(lib.hb_fv_set_fixed_boundary(self.fv, i-1, U, self.eqn.numStates), 'hb_fv_set_fixed_boundary')
		end
	end
end

-- solver.ops (euler.lua:179-188, mhd.lua:110-122): the ops this library builds are registered with the fused path, which then runs
-- op:addSource inside every stage and op:step after the integrator (solverbase.lua:3219-3237); call after refreshSolverProgram
function B200Solver:registerOps()
	for _, op in ipairs(self.ops) do
		local d = ffi.new'hb_op_desc'
		if op.name == 'selfgrav' then
			if self.useGravity then
				d.kind = lib.HB_OP_SELFGRAV
				d.param = self.eqn.guiVars.gravitationalConstant.value      -- / unit_m3_per_kg_s2 = 1 in the default units
			end
		elseif op.name == 'NoDiv' then
			assert(require 'hydro.op.poisson_jacobi':isa(op), "hydrob200: NoDiv needs noDivPoissonSolver=jacobi (the krylov parent is not built)")
			d.kind = lib.HB_OP_NODIV
		else
			error("hydrob200: op not built: "..tostring(op.name))
		end
		if d.kind ~= 0 then
			d.max_iters, d.stop_on_epsilon, d.stop_epsilon = op.maxIters, op.stopOnEpsilon and 1 or 0, op.stopEpsilon
			check(lib.hb_fv_add_op(self.fv, d, nil), 'hb_fv_add_op')
		end
	end
end
-- the op:resetState() loop of SolverBase:resetState (solverbase.lua:2106-2111)
function B200Solver:resetOps() check(lib.hb_fv_ops_reset(self.fv), 'hb_fv_ops_reset') end

function B200Solver:uploadState(aosPtr) check(lib.hb_fv_set_state(self.fv, aosPtr), 'hb_fv_set_state') end
function B200Solver:downloadState(aosPtr) check(lib.hb_fv_get_state(self.fv, aosPtr), 'hb_fv_get_state') end

function B200Solver:boundary() check(lib.hb_fv_boundary(self.fv), 'hb_fv_boundary') end
function B200Solver:constrainU() check(lib.hb_fv_constrainU(self.fv), 'hb_fv_constrainU') end
function B200Solver:calcDT()
	local dt = ffi.new'double[1]'
	check(lib.hb_fv_calc_dt(self.fv, dt), 'hb_fv_calc_dt')
	return dt[0]
end
function B200Solver:step(dt) check(lib.hb_fv_step(self.fv, dt), 'hb_fv_step') end
function B200Solver:update()
	check(lib.hb_fv_update(self.fv, 1), 'hb_fv_update')
	local t, dt = ffi.new'double[1]', ffi.new'double[1]'
	check(lib.hb_fv_get_time(self.fv, t, dt), 'hb_fv_get_time')
	self.t, self.dt = t[0], dt[0]
end
function B200Solver:calcDeriv(derivHostPtr, dt) check(lib.hb_fv_calc_deriv(self.fv, dt, derivHostPtr), 'hb_fv_calc_deriv') end

return B200Solver
