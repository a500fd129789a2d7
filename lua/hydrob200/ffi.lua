--[[
hydrob200/ffi.lua -- LuaJIT FFI binding of libhydrob200.so (include/hydrob200.h).

This is the reference-side stub a hydro-cl-lua maintainer adds: it replaces `require 'cl'` / `require 'cl.obj.*'`
for the finite-volume path.  It declares exactly the entry points of include/hydrob200.h (the header is read at run
time, so the two cannot drift) and wraps the int status into Lua errors, as lua-opencl's `classert` does.

NOT executed in the authoring container (no LuaJIT there); the same C ABI is exercised by the Python/ctypes binding
in hydro-cl-lua_b200/_lib.py and by tests/test_abi.py.
--]]
local ffi = require 'ffi'

local function readHeader(path)
	local f = assert(io.open(path, 'rb'), "cannot open "..path)
	local src = f:read'*a'
	f:close()
	-- ffi.cdef does not run a preprocessor: drop directives, turn the #define constants into an enum
	local enums = {}
	src = src:gsub('#define%s+(HB_[%w_]+)%s+(%-?%d+)[^\n]*', function(k, v)
		enums[#enums+1] = ('%s = %s'):format(k, v)
		return ''
	end)
	src = src:gsub('#[^\n]*', ''):gsub('extern%s+"C"%s*{', ''):gsub('\n}%s*\n', '\n')
	return 'enum { '..table.concat(enums, ', ')..' };\n'..src
end

local root = os.getenv'HYDROB200_ROOT' or '.'
ffi.cdef(readHeader(root..'/include/hydrob200.h'))
local lib = ffi.load(root..'/hydro-cl-lua_b200/csrc/libhydrob200.so')

local M = {lib = lib, ffi = ffi}

function M.check(code, what)
	if code ~= 0 then
		error(('%s: libhydrob200 error %d: %s'):format(what or 'hydrob200', code, ffi.string(lib.hb_last_error())), 2)
	end
end

return M
