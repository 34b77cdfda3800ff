// jit.hpp -- optional run-time specialisation of the sampling kernels (host side).
//
// The reference evaluates a rule form by `eval` -- SBCL compiles it -- every time it is tested
// (good-opening?, bidding.cl:60-74).  The bytecode interpreter in features.cuh is the always-
// available device equivalent; for long jobs the library can instead translate the constraint's rule
// forms to C++ (emit_rule below), splice them into the SAME hand-written kernel text (kernels.cuh,
// embedded at build time) and compile that with NVRTC for sm_100a.  Results are bit-identical to the
// interpreter path (tests/test_gpu_rules.py); only the rule evaluation gets cheaper.
// libnvrtc and the CUDA driver API are bound with dlopen, so a missing NVRTC only disables the JIT.
#pragma once
#include <cuda.h>
#include <dlfcn.h>
#include <nvrtc.h>

#include <mutex>
#include <string>
#include <vector>

#include "rules.hpp"

namespace bmc {

// ---- S-expression -> C++ expression (same language as RuleCompiler::expr) -----------------
class RuleEmitter {
public:
    std::string emit(const Sx &x) { return expr(x); }

private:
    static const char *var_of(const std::string &s)
    {
        static const char *in[] = {"C", "D", "H", "S", "CPOWER", "DPOWER", "HPOWER", "SPOWER", "HCP", "LOOSERS", "LOSERS"};
        static const char *out[] = {"vC", "vD", "vH", "vS", "vCP", "vDP", "vHP", "vSP", "vHCP", "vLOS", "vLOS"};
        for (int i = 0; i < 11; ++i) if (s == in[i]) return out[i];
        return nullptr;
    }
    static bool const_value(const Sx &x, long &v)
    {
        if (x.kind == Sx::INT) { v = x.num; return true; }
        if (x.is_list() && !x.items.empty() && x.items[0].kind == Sx::SYM &&
            (x.items[0].sym == "+" || x.items[0].sym == "-" || x.items[0].sym == "*")) {
            const char op = x.items[0].sym[0];
            long acc = op == '*' ? 1 : 0;
            for (size_t i = 1; i < x.items.size(); ++i) {
                long t;
                if (!const_value(x.items[i], t)) return false;
                if (op == '+') acc += t;
                else if (op == '*') acc *= t;
                else acc = (i == 1 && x.items.size() > 2) ? t : acc - t;
            }
            v = acc;
            return true;
        }
        return false;
    }
    std::string operand(const Sx &x)
    {
        long v;
        if (const_value(x, v)) return std::to_string(v);
        if (x.kind == Sx::SYM) {
            const char *n = var_of(x.sym);
            if (n) return n;
        }
        throw RuleError("operand is neither an integer nor a hand variable");
    }
    static const char *cmp_of(const std::string &s)
    {
        if (s == "<") return "<";
        if (s == "<=") return "<=";
        if (s == ">") return ">";
        if (s == ">=") return ">=";
        if (s == "=") return "==";
        if (s == "/=") return "!=";
        return nullptr;
    }
    std::string expr(const Sx &x)
    {
        switch (x.kind) {
        case Sx::NIL: return "false";
        case Sx::INT: case Sx::STR: return "true";
        case Sx::SYM:
            if (x.sym == "T") return "true";
            if (x.sym == "BALANCED") return "vBAL";
            if (var_of(x.sym)) return "true";
            throw RuleError("unknown variable: " + x.sym);
        case Sx::LIST: break;
        }
        const Sx &h = x.items[0];
        // an "illegal function call" (bidding.cl:249) counts only where C++'s own short-circuit order -- the
        // same as Lisp's -- reaches it
        if (h.kind != Sx::SYM) return "(err = true, false)";
        const std::string &op = h.sym;
        const size_t n = x.items.size() - 1;
        if (op == "AND" || op == "OR") {
            const bool is_and = op == "AND";
            if (n == 0) return is_and ? "true" : "false";
            std::string s = "(";
            for (size_t i = 1; i <= n; ++i) s += (i > 1 ? (is_and ? " && " : " || ") : "") + expr(x.items[i]);
            return s + ")";
        }
        if (op == "NOT" || op == "NULL") return "(!" + expr(x.items.at(1)) + ")";
        if (const char *c = cmp_of(op)) {
            if (n == 1) { operand(x.items[1]); return "true"; }
            std::string s = "(";
            for (size_t i = 1; i < n; ++i)
                s += std::string(i > 1 ? " && " : "") + "(" + operand(x.items[i]) + " " + c + " " + operand(x.items[i + 1]) + ")";
            return s + ")";
        }
        if (op == "COND") return cond_from(x, 1);
        if (op == "IF") return "(" + expr(x.items.at(1)) + " ? " + expr(x.items.at(2)) + " : " + (n == 3 ? expr(x.items[3]) : "false") + ")";
        if (op == "FIND-IF") return find_if(x);
        if (op == "QUOTE") return (x.items.size() > 1 && x.items[1].kind != Sx::NIL) ? "true" : "false";
        throw RuleError("form outside the bidding.cl predicate language");
    }
    std::string cond_from(const Sx &x, size_t i)
    {
        if (i >= x.items.size()) return "false";
        const Sx &cl = x.items[i];
        const std::string t = expr(cl.items.at(0));
        const std::string v = cl.items.size() > 1 ? expr(cl.items.back()) : "true";
        return "(" + t + " ? " + v + " : " + cond_from(x, i + 1) + ")";
    }
    std::string find_if(const Sx &x)
    {
        const Sx &fn = x.items.at(1);
        const char *c = cmp_of(fn.items.at(1).items.at(1).sym);
        if (!c) throw RuleError("unsupported find-if predicate");
        const std::string nn = operand(fn.items.at(2));
        const Sx &seq = x.items.at(2);
        static const char *lens[] = {"vC", "vD", "vH", "vS"}, *power[] = {"vCP", "vDP", "vHP", "vSP"};
        const char **base = nullptr;
        int from = 0, to = 4;
        if (seq.is_sym("LENS")) base = lens;
        else if (seq.is_sym("POWER")) base = power;
        else {
            base = seq.items.at(1).is_sym("LENS") ? lens : power;
            long v;
            const_value(seq.items.at(2), v); from = (int)v;
            if (seq.items.size() == 4) { const_value(seq.items[3], v); to = (int)v; }
        }
        if (from >= to) return "false";
        std::string s = "(";
        for (int i = from; i < to; ++i) s += std::string(i > from ? " || " : "") + "(" + nn + " " + c + " " + base[i] + ")";
        return s + ")";
    }
};

// ---- dynamic binding of NVRTC and the driver API --------------------------------------------
struct JitApi {
    bool ok = false;
    std::string why;
    decltype(&nvrtcCreateProgram) CreateProgram;
    decltype(&nvrtcCompileProgram) CompileProgram;
    decltype(&nvrtcGetProgramLogSize) GetProgramLogSize;
    decltype(&nvrtcGetProgramLog) GetProgramLog;
    decltype(&nvrtcGetCUBINSize) GetCUBINSize;
    decltype(&nvrtcGetCUBIN) GetCUBIN;
    decltype(&nvrtcDestroyProgram) DestroyProgram;
    decltype(&cuModuleLoadData) ModuleLoadData;
    decltype(&cuModuleUnload) ModuleUnload;
    decltype(&cuModuleGetFunction) ModuleGetFunction;
    decltype(&cuModuleGetGlobal) ModuleGetGlobal;
    decltype(&cuMemcpyHtoD) MemcpyHtoD;
    decltype(&cuLaunchKernel) LaunchKernel;
    decltype(&cuOccupancyMaxActiveBlocksPerMultiprocessor) Occupancy;

    static JitApi &get()
    {
        static JitApi api;
        static std::once_flag once;
        std::call_once(once, [] { api.load(); });
        return api;
    }

private:
    void load()
    {
        void *rtc = nullptr;
        for (const char *n : {"libnvrtc.so.12", "libnvrtc.so", "/usr/local/cuda/lib64/libnvrtc.so.12"})
            if ((rtc = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        void *drv = nullptr;
        for (const char *n : {"libcuda.so.1", "libcuda.so"})
            if ((drv = dlopen(n, RTLD_NOW | RTLD_GLOBAL))) break;
        if (!rtc || !drv) { why = std::string("cannot load ") + (!rtc ? "libnvrtc" : "libcuda"); return; }
        bool all = true;
        auto sym = [&](void *lib, const char *name, auto &fn) {
            void *p = dlsym(lib, name);
            if (!p) { all = false; why = std::string("missing symbol ") + name; }
            fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(p);
        };
        sym(rtc, "nvrtcCreateProgram", CreateProgram);
        sym(rtc, "nvrtcCompileProgram", CompileProgram);
        sym(rtc, "nvrtcGetProgramLogSize", GetProgramLogSize);
        sym(rtc, "nvrtcGetProgramLog", GetProgramLog);
        sym(rtc, "nvrtcGetCUBINSize", GetCUBINSize);
        sym(rtc, "nvrtcGetCUBIN", GetCUBIN);
        sym(rtc, "nvrtcDestroyProgram", DestroyProgram);
        sym(drv, "cuModuleLoadData", ModuleLoadData);
        sym(drv, "cuModuleUnload", ModuleUnload);
        sym(drv, "cuModuleGetFunction", ModuleGetFunction);
        sym(drv, "cuModuleGetGlobal_v2", ModuleGetGlobal);
        sym(drv, "cuMemcpyHtoD_v2", MemcpyHtoD);
        sym(drv, "cuLaunchKernel", LaunchKernel);
        sym(drv, "cuOccupancyMaxActiveBlocksPerMultiprocessor", Occupancy);
        ok = all;
    }
};

struct JitKernel {
    CUmodule module = nullptr;
    CUfunction func = nullptr;
    int device = -1;
    int blocks_per_sm = 1;
    float compile_ms = 0.f;
};

}  // namespace bmc
