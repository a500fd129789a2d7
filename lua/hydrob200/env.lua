--[[
hydrob200/env.lua -- the subset of lua-opencl's `cl.obj.env` / `cl.obj.buffer` / `cl.obj.program` / `cl.obj.kernel` /
`cl.obj.reduce` that hydro/solver/*.lua and hydro/int/*.lua use (SURVEY.md 8b), on top of the C ABI.

	hydro/app.lua:891-929          CLEnv{precision=...}            -> Env{precision=..., device=...}
	hydro/solver/solverbase.lua:1077-1095,1417-1428  env:buffer / CLBuffer:fromCPU/toCPU/fill   -> Buffer
	hydro/solver/solverbase.lua:558-713,1696-1699    Program{code=...}:compile()               -> Program (NVRTC, sm_100a)
	hydro/solver/fvsolver.lua:216-221, solverbase.lua:1328-1345  program:kernel(...), k.obj:setArg, k(...) -> Kernel
	hydro/solver/solverbase.lua:1350-1376            env:reduce{op=...}                          -> reduce closure

Kernel source handed to Program is the OpenCL-C the reference's templates emit.  hb_module_compile_opencl (csrc/hb_cl.cu) rewrites
the dialect on C tokens -- `kernel` / `global` / `constant` qualifiers, `(real3){.x=..}` compound literals (also inside #define
bodies, hydro/code/math.cl:47-52) -- prepends hb_cl_prelude() (get_global_id & co., `typedef HB_REAL real`) and compiles with NVRTC
-default-device for sm_100a, so the per-equation device functions of hydro/eqn/*.cl compile without a hand conversion.  (Round 1 tried
this with a macro prelude in this file: `#define global` also erases the token inside CUDA's own `__global__`, and no macro turns a
compound literal into C++.)  tests/test_gpu_fine_grained.py drives exactly this surface from Python on the GPU, including one whole
unfused RK4 update that is bit-identical to hb_fv_update.
--]]
local hb = require 'hydrob200.ffi'
local ffi, lib, check = hb.ffi, hb.lib, hb.check
local class = require 'ext.class'

local Env = class()
function Env:init(args)
	args = args or {}
	self.real = args.precision == 'float' and 'float' or 'double'       -- hydro/app.lua:892
	local p = ffi.new'hb_ctx*[1]'
	check(lib.hb_ctx_create(args.device or 0, self.real == 'float' and 4 or 8, p), 'hb_ctx_create')
	self.ctx = ffi.gc(p[0], lib.hb_ctx_destroy)
	self.cmds = {self}                                                   -- env.cmds[1]:finish()
	self.code = ''                                                       -- (the prelude is added by hb_module_compile_opencl)
end
function Env:finish() check(lib.hb_sync(self.ctx), 'hb_sync') end
function Env:buffer(args) return require'hydrob200.env'.Buffer(self, args) end
function Env:program(args) return require'hydrob200.env'.Program(self, args) end
function Env:reduce(args)
	local ops = {min = lib.HB_REDUCE_MIN, max = lib.HB_REDUCE_MAX, sum = lib.HB_REDUCE_SUM}
	local out = ffi.new'double[1]'
	return function(buf, count)
		buf = buf or args.buffer
		check(lib.hb_reduce(self.ctx, buf.obj or buf, count or args.count, ops[args.op or 'min'], out), 'hb_reduce')
		return out[0]
	end
end

local Buffer = class()
function Buffer:init(env, args)
	self.env, self.type, self.count = env, args.type or env.real, args.count or args.size
	self.size = self.count * ffi.sizeof(self.type)
	local p = ffi.new'hb_buf*[1]'
	check(lib.hb_buf_alloc(env.ctx, self.size, p), 'hb_buf_alloc')
	self.obj = ffi.gc(p[0], lib.hb_buf_free)
	if args.data then self:fromCPU(args.data) end
end
function Buffer:fromCPU(ptr) check(lib.hb_buf_write(self.obj, ptr, 0, self.size), 'hb_buf_write'); return self end
function Buffer:toCPU(ptr)
	ptr = ptr or ffi.new(self.type..'[?]', self.count)
	check(lib.hb_buf_read(self.obj, ptr, 0, self.size), 'hb_buf_read')
	return ptr
end
function Buffer:fill(value)
	local v = ffi.new(self.env.real..'[1]', value or 0)
	check(lib.hb_buf_fill(self.obj, v, ffi.sizeof(self.env.real), 0, self.size), 'hb_buf_fill')
end

local Kernel = class()
function Kernel:init(program, args)
	self.program = program
	local p = ffi.new'hb_kernel*[1]'
	check(lib.hb_kernel_get(program.obj, args.name, p), 'hb_kernel_get')
	self.obj = self                                                     -- k.obj:setArg(i, x)
	self.h = ffi.gc(p[0], lib.hb_kernel_free)
	self.domain = args.domain
	if args.setArgs then for i, a in ipairs(args.setArgs) do self:setArg(i-1, a) end end
end
function Kernel:setArg(i, x)
	if type(x) == 'table' and x.obj then x = x.obj end
	if ffi.istype('hb_buf*', x) then check(lib.hb_kernel_set_arg_buf(self.h, i, x), 'hb_kernel_set_arg_buf')
	elseif type(x) == 'number' then
		local v = ffi.new(self.program.env.real..'[1]', x)
		check(lib.hb_kernel_set_arg(self.h, i, v, ffi.sizeof(v)), 'hb_kernel_set_arg')
	else check(lib.hb_kernel_set_arg(self.h, i, x, ffi.sizeof(x)), 'hb_kernel_set_arg') end
end
function Kernel:__call(...)
	for i = 1, select('#', ...) do self:setArg(i-1, (select(i, ...))) end
	local d = self.domain
	check(lib.hb_kernel_launch(self.h, ffi.new('size_t[3]', d.globalSize), ffi.new('size_t[3]', d.localSize), 0), 'hb_kernel_launch')
end

local Program = class()
function Program:init(env, args) self.env, self.code, self.name = env, args.code, args.name or 'program' end
function Program:compile(args)
	local p = ffi.new'hb_module*[1]'
	local log = ffi.new('char[?]', 1 << 16)
	local rc = lib.hb_module_compile_opencl(self.env.ctx, self.env.code..self.code, self.name, nil, 0, p, log, 1 << 16)
	self.log = ffi.string(log)
	if rc ~= 0 then error(self.name..': NVRTC build failed:\n'..self.log) end
	self.obj = ffi.gc(p[0], lib.hb_module_free)
	return self
end
function Program:kernel(args, ...)
	if type(args) == 'string' then args = {name = args, setArgs = {...}} end
	return Kernel(self, args)
end

return {Env = Env, Buffer = Buffer, Program = Program, Kernel = Kernel}
