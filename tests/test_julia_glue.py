"""The Julia glue cannot be executed here (no Julia in the image), so at least its FFI signatures are checked mechanically: every
`ccall((:dmt_xxx, libdmt), RetType, (ArgTypes...), args...)` in julia/DiffusionMCMCToolsB200.jl must name an entry point that
include/dmt.h declares, with the same number of arguments, compatible argument types and the declared return type, and must pass
exactly as many values as it declares types."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

C2JL = {
    "dmt_ctx *": {"Ptr{Cvoid}"}, "const dmt_ctx *": {"Ptr{Cvoid}"}, "dmt_ctx **": {"Ref{Ptr{Cvoid}}"},
    "const dmt_config *": {"Ref{DmtConfig}"},
    "int32_t": {"Int32"}, "uint32_t": {"UInt32"}, "uint64_t": {"UInt64"}, "double": {"Float64"},
    "double *": {"Ptr{Float64}", "Ref{Float64}"}, "const double *": {"Ptr{Float64}"},
    "int32_t *": {"Ptr{Int32}", "Ref{Int32}"}, "const int32_t *": {"Ptr{Int32}"},
    "uint8_t *": {"Ptr{UInt8}"}, "const uint8_t *": {"Ptr{UInt8}"},
    "int64_t *": {"Ptr{Int64}"}, "uint64_t *": {"Ptr{UInt64}", "Ref{UInt64}"}, "char *": {"Ptr{UInt8}"}, "void **": {"Ref{Ptr{Cvoid}}"},
}
RET2JL = {"int32_t": "Int32", "const char *": "Cstring"}


def header_prototypes():
    txt = open(os.path.join(ROOT, "include", "dmt.h")).read()
    txt = re.sub(r"/\*.*?\*/", " ", txt, flags=re.S)
    protos = {}
    for m in re.finditer(r"(int32_t|const char \*)\s*(dmt_\w+)\s*\(([^;{]*?)\)\s*;", txt, flags=re.S):
        ret, name, args = m.group(1).strip(), m.group(2), " ".join(m.group(3).split())
        types = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                t = re.sub(r"\b\w+$", "", a).strip() if not a.endswith("*") else a      # drop the parameter name
                t = re.sub(r"\s*\*", " *", t)
                t = re.sub(r"\* \*", "**", t)
                types.append(" ".join(t.split()))
        protos[name] = (ret, types)
    return protos


def split_top(s):
    """split on commas that are not nested in (), {}, []"""
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "({[":
            depth += 1
        elif ch in ")}]":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def julia_ccalls():
    txt = open(os.path.join(ROOT, "julia", "DiffusionMCMCToolsB200.jl")).read()
    calls = []
    for m in re.finditer(r"ccall\(\(:(dmt_\w+), libdmt\),", txt):
        i, depth = m.start() + len("ccall"), 0
        j = i
        while True:                              # the matching parenthesis of ccall(
            if txt[j] == "(":
                depth += 1
            elif txt[j] == ")":
                depth -= 1
                if depth == 0:
                    break
            j += 1
        parts = split_top(txt[i + 1:j])
        ret, argt, vals = parts[1], parts[2], parts[3:]
        assert argt.startswith("(") and argt.endswith(")"), (m.group(1), argt)
        types = split_top(argt[1:-1].rstrip(","))
        calls.append((m.group(1), ret, types, vals, txt.count("\n", 0, m.start()) + 1))
    return calls


def test_every_ccall_matches_the_header():
    protos = header_prototypes()
    calls = julia_ccalls()
    assert len(protos) >= 60 and len(calls) >= 40
    for name, ret, types, vals, line in calls:
        where = "julia/DiffusionMCMCToolsB200.jl:%d %s" % (line, name)
        assert name in protos, where + ": not declared in include/dmt.h"
        cret, ctypes_ = protos[name]
        assert RET2JL[cret] == ret, (where, ret, cret)
        assert len(types) == len(ctypes_), (where, types, ctypes_)
        assert len(vals) == len(types), (where, "passes %d values for %d declared types" % (len(vals), len(types)))
        for k, (jt, ct) in enumerate(zip(types, ctypes_)):
            assert ct in C2JL, (where, "unmapped C type", ct)
            assert jt in C2JL[ct], (where, "argument %d: Julia %s vs C %s" % (k, jt, ct))


def test_glue_covers_the_reference_exports():
    """every generic function DiffusionMCMCTools exports for BlockEnsemble-level work has a device method in the glue"""
    txt = open(os.path.join(ROOT, "julia", "DiffusionMCMCToolsB200.jl")).read()
    for f in ["draw_proposal_path!", "accept_reject_proposal_path!", "swap_paths!", "swap_XX!", "swap_WW!", "swap_PP!", "swap_ll!", "loglikhd!",
              "loglikhd°!", "fetch_ll", "fetch_ll°", "save_ll!", "accpt_rate", "ll_of_accepted", "find_W_for_X!", "set_proposal_law!",
              "set_accepted!", "set_ll!", "recompute_path!", "GP.set_obs!", "GP.recompute_guiding_term!"]:
        assert re.search(r"(^|\n|\(:)\s*(function\s+)?%s\(|:%s," % (re.escape(f), re.escape(f)), txt), f
