// hb_cl.cu -- OpenCL-C dialect front end of the program API: lets the kernel source the reference's Lua templates emit
// (hydro/code/math.cl, hydro/eqn/cl/*.cl, hydro/solver/*.cl, hydro/eqn/*.cl after template expansion) be handed to
// hb_module_compile_opencl() instead of a hand-converted CUDA file.  SURVEY 2.5 lists the OpenCL-isms of those templates:
//
//   `kernel`, `global`, `constant` address-space / function qualifiers         -> token rewrite below
//   `(real3){.x = a, .y = b, .z = c}` C99 compound literals with designators   -> `([&]{ real3 _v{}; _v.x = a; ... return _v; }())`
//        (the reference builds every vector / tensor value this way: hydro/code/math.cl:47-52 `_real3(a,b,c)`; its vector types are
//        unions of anonymous structs, for which C++20 designated initialisers are not valid: "duplicate designator")
//   `(T){a, b, c}` compound literals without designators                       -> `T{a, b, c}`
//   `real3 .s0 / .s[i]` swizzles, `int4`                                       -> nothing to do: the emitted typedefs carry them (math.cl:27-33),
//        int4 is a CUDA built-in with x, y, z, w
//   get_global_id / get_local_id / get_group_id / get_global_size / barrier    -> prelude (hb_cl_prelude)
//   unannotated `static inline` device functions                               -> NVRTC `-default-device`
//
// A macro prelude alone cannot do this (round 1's lua/hydrob200/env.lua tried): `#define global` also erases the token inside
// CUDA's own `__global__` = `__location__(global)`, and no macro can turn a compound literal into C++.  The rewrite works on C
// tokens (identifiers, numbers, strings, comments, punctuation), never inside comments, strings or preprocessor lines.
#include "hb_core.h"
#include <cctype>
#include <cstring>
#include <string>
#include <vector>

namespace hb {

struct Tok { int kind; std::string text; };   // kind: 0 space/comment/preprocessor, 1 identifier, 2 number, 3 string/char, 4 punctuation

static std::vector<Tok> tokenize(const std::string& s) {
	std::vector<Tok> out;
	size_t i = 0, n = s.size();
	bool lineStart = true;
	while (i < n) {
		char c = s[i];
		size_t j = i;
		if (c == '/' && i + 1 < n && s[i + 1] == '/') { while (j < n && s[j] != '\n') ++j; out.push_back({0, s.substr(i, j - i)}); }
		else if (c == '/' && i + 1 < n && s[i + 1] == '*') { j = s.find("*/", i + 2); j = j == std::string::npos ? n : j + 2; out.push_back({0, s.substr(i, j - i)}); }
		else if (c == '#' && lineStart) {   // preprocessor line (with continuations)
			while (j < n && !(s[j] == '\n' && (j == 0 || s[j - 1] != '\\'))) ++j;
			std::string const line = s.substr(i, j - i);
			size_t k = 1;
			while (k < line.size() && (line[k] == ' ' || line[k] == '\t')) ++k;
			if (line.compare(k, 6, "define") == 0) {
				// a macro body is code (hydro/code/math.cl:47-52: `#define _real3(a,b,c) ((real3){.x=a, .y=b, .z=c})`): rewrite it like code
				out.push_back({0, line.substr(0, k + 6)});
				std::vector<Tok> body = tokenize(line.substr(k + 6));
				for (auto& b : body) out.push_back(b);
			}
			else out.push_back({0, line});              // #include, #if, #pragma ...: passed through untouched
		}
		else if (isspace((unsigned char)c)) { while (j < n && isspace((unsigned char)s[j])) { if (s[j] == '\n') lineStart = true; ++j; } out.push_back({0, s.substr(i, j - i)}); i = j; continue; }
		else if (isalpha((unsigned char)c) || c == '_') { while (j < n && (isalnum((unsigned char)s[j]) || s[j] == '_')) ++j; out.push_back({1, s.substr(i, j - i)}); }
		else if (isdigit((unsigned char)c) || (c == '.' && i + 1 < n && isdigit((unsigned char)s[i + 1]))) {
			while (j < n && (isalnum((unsigned char)s[j]) || s[j] == '.' || ((s[j] == '+' || s[j] == '-') && (s[j - 1] == 'e' || s[j - 1] == 'E')))) ++j;
			out.push_back({2, s.substr(i, j - i)});
		}
		else if (c == '"' || c == '\'') { ++j; while (j < n && s[j] != c) { if (s[j] == '\\') ++j; ++j; } j = j < n ? j + 1 : n; out.push_back({3, s.substr(i, j - i)}); }
		else { j = i + 1; out.push_back({4, std::string(1, c)}); }
		lineStart = false;
		i = j;
	}
	return out;
}

static int nextSig(const std::vector<Tok>& t, int i) { for (++i; i < (int)t.size(); ++i) if (t[i].kind) return i; return -1; }
static int prevSig(const std::vector<Tok>& t, int i) { for (--i; i >= 0; --i) if (t[i].kind) return i; return -1; }
static bool isP(const Tok& t, char c) { return t.kind == 4 && t.text[0] == c; }

// index of the token closing the bracket opened at t[open]
static int matchClose(const std::vector<Tok>& t, int open) {
	char const o = t[open].text[0], c = o == '(' ? ')' : (o == '{' ? '}' : ']');
	int depth = 0;
	for (int i = open; i < (int)t.size(); ++i) {
		if (isP(t[i], o)) ++depth;
		else if (isP(t[i], c) && --depth == 0) return i;
	}
	return -1;
}

static std::string join(const std::vector<Tok>& t, int a, int b) { std::string s; for (int i = a; i < b; ++i) s += t[i].text; return s; }

static std::string translateTokens(std::vector<Tok> t, int depthGuard = 0);

// `( T ) { list }` at t[lp] ... : returns the replacement text and sets `end` to the index after the closing brace, or returns "" when
// this is not a compound literal
static std::string compoundLiteral(const std::vector<Tok>& t, int lp, int& end, int depthGuard) {
	int const id = nextSig(t, lp);
	if (id < 0 || t[id].kind != 1) return "";
	int const rp = nextSig(t, id);
	if (rp < 0 || !isP(t[rp], ')')) return "";
	int const lb = nextSig(t, rp);
	if (lb < 0 || !isP(t[lb], '{')) return "";
	int const p = prevSig(t, lp);
	if (p >= 0) {
		const Tok& q = t[p];
		if (q.kind == 1 && q.text != "return") return "";                 // f(x) { ... : a function definition, if / for / while / switch (x) {
		if (q.kind == 2 || q.kind == 3 || isP(q, ')') || isP(q, ']')) return "";
	}
	int const rb = matchClose(t, lb);
	if (rb < 0) return "";
	// split the initialiser list at top-level commas
	std::vector<std::pair<int, int>> items;
	int depth = 0, start = lb + 1;
	for (int i = lb + 1; i < rb; ++i) {
		if (isP(t[i], '(') || isP(t[i], '{') || isP(t[i], '[')) ++depth;
		else if (isP(t[i], ')') || isP(t[i], '}') || isP(t[i], ']')) --depth;
		else if (isP(t[i], ',') && depth == 0) { items.push_back({start, i}); start = i + 1; }
	}
	if (nextSig(t, start - 1) >= 0 && nextSig(t, start - 1) < rb) items.push_back({start, rb});
	bool designated = false;
	for (auto& it : items) { int const f = nextSig(t, it.first - 1); if (f >= 0 && f < it.second && isP(t[f], '.')) designated = true; }
	std::string const T = t[id].text;
	std::string out;
	auto sub = [&](int a, int b) { return translateTokens(std::vector<Tok>(t.begin() + a, t.begin() + b), depthGuard + 1); };
	if (!designated) {
		out = T + "{";
		for (size_t k = 0; k < items.size(); ++k) out += (k ? "," : "") + sub(items[k].first, items[k].second);
		out += "}";
	} else {
		out = "([&]{ " + T + " _hb_v{}; ";
		for (auto& it : items) {
			int const f = nextSig(t, it.first - 1);
			if (!(f >= 0 && f < it.second && isP(t[f], '.'))) return "";   // mixed positional / designated: leave it to the compiler's error message
			out += "_hb_v" + sub(f, it.second) + "; ";
		}
		out += "return _hb_v; }())";
	}
	end = rb + 1;
	return out;
}

static std::string translateTokens(std::vector<Tok> t, int depthGuard) {
	std::string out;
	int paren = 0, brace = 0;
	for (int i = 0; i < (int)t.size(); ++i) {
		const Tok& k = t[i];
		if (k.kind == 4) {
			if (k.text[0] == '(') {
				if (depthGuard < 32) {
					int end = 0;
					std::string const r = compoundLiteral(t, i, end, depthGuard);
					if (!r.empty()) { out += r; i = end - 1; continue; }
				}
				++paren;
			}
			else if (k.text[0] == ')') --paren;
			else if (k.text[0] == '{') ++brace;
			else if (k.text[0] == '}') --brace;
		}
		if (k.kind == 1) {
			const std::string& w = k.text;
			if (w == "kernel" || w == "__kernel") { out += "extern \"C\" __global__"; continue; }
			if (w == "global" || w == "__global") continue;                                    // CUDA has one global address space: no qualifier
			if (w == "constant" || w == "__constant") {
				// a pointer-parameter / local qualifier is dropped (the pointee's own `const` stays); a file-scope object becomes __constant__
				if (paren == 0 && brace == 0) out += "__constant__ const";
				continue;
			}
			if (w == "restrict") { out += "__restrict__"; continue; }
		}
		out += k.text;
	}
	return out;
}

static const char* kPrelude =
	"// hb_cl_prelude: OpenCL-C built-ins of the reference's kernel templates on CUDA (hb_cl.cu)\n"
	"#define get_global_id(i) ((int)((i) == 0 ? blockIdx.x * blockDim.x + threadIdx.x : (i) == 1 ? blockIdx.y * blockDim.y + threadIdx.y : blockIdx.z * blockDim.z + threadIdx.z))\n"
	"#define get_local_id(i) ((int)((i) == 0 ? threadIdx.x : (i) == 1 ? threadIdx.y : threadIdx.z))\n"
	"#define get_group_id(i) ((int)((i) == 0 ? blockIdx.x : (i) == 1 ? blockIdx.y : blockIdx.z))\n"
	"#define get_global_size(i) ((int)((i) == 0 ? gridDim.x * blockDim.x : (i) == 1 ? gridDim.y * blockDim.y : gridDim.z * blockDim.z))\n"
	"#define get_local_size(i) ((int)((i) == 0 ? blockDim.x : (i) == 1 ? blockDim.y : blockDim.z))\n"
	"#define get_num_groups(i) ((int)((i) == 0 ? gridDim.x : (i) == 1 ? gridDim.y : gridDim.z))\n"
	"#define CLK_LOCAL_MEM_FENCE 1\n"
	"#define CLK_GLOBAL_MEM_FENCE 2\n"
	"#define barrier(flags) __syncthreads()\n"
	"#define mem_fence(flags) __threadfence()\n"
	"#ifndef INFINITY\n#define INFINITY (__longlong_as_double(0x7ff0000000000000LL))\n#endif\n"
	"#ifndef M_PI\n#define M_PI 3.14159265358979323846\n#endif\n"
	"typedef unsigned int uint;\n"
	"typedef unsigned long ulong;\n";

}   // namespace hb

using namespace hb;

extern "C" {

const char* hb_cl_prelude(void) { return kPrelude; }

// OpenCL-C dialect -> CUDA C++ (the rewrite only; no prelude).  Returns HB_OK and writes a NUL-terminated string when it fits `cap`;
// *needed (optional) receives the size including the terminator either way.
int hb_cl_translate(const char* src, char* out, size_t cap, size_t* needed) {
	if (!src) return setError(HB_ERR_INVALID, "hb_cl_translate: null source");
	std::string const r = translateTokens(tokenize(src));
	if (needed) *needed = r.size() + 1;
	if (!out || cap < r.size() + 1) return out ? setError(HB_ERR_INVALID, "hb_cl_translate: output buffer too small") : HB_OK;
	memcpy(out, r.c_str(), r.size() + 1);
	return HB_OK;
}

// Program{name, code}:compile() for OpenCL-C dialect source (hydro/solver/solverbase.lua:558-713,1686-1699): prelude + rewrite, then
// NVRTC with -default-device (the templates' `static inline` functions carry no execution-space annotation) and C++20.
int hb_module_compile_opencl(hb_ctx* ctx, const char* cl_src, const char* name, const char* const* opts, int nopts,
	hb_module** out, char* log, size_t log_cap)
{
	if (!cl_src) return setError(HB_ERR_INVALID, "hb_module_compile_opencl: null source");
	std::string const cu = std::string(kPrelude) + "typedef HB_REAL real;\n" + translateTokens(tokenize(cl_src));
	std::vector<const char*> o;
	o.push_back("-default-device");
	o.push_back("--std=c++20");
	for (int i = 0; i < nopts; ++i) if (opts && opts[i]) o.push_back(opts[i]);
	return hb_module_compile(ctx, cu.c_str(), name, o.data(), (int)o.size(), out, log, log_cap);
}

}   // extern "C"
