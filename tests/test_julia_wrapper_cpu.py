"""julia/RLSTDPB200.jl cannot be executed here (no Julia): check statically that its struct mirrors have the
fields of include/snn_b200.h in the same order with matching types, and that every symbol it ccalls exists."""
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CTYPE = {"int32_t": "Int32", "int64_t": "Int64", "double": "Float64", "uint64_t": "UInt64", "uint8_t": "UInt8"}


def _strip_comments(text):
    return re.sub(r"/\*.*?\*/", "", text, flags=re.S)


def c_struct(name):
    hdr = _strip_comments(open(os.path.join(ROOT, "include", "snn_b200.h")).read())
    body = re.search(r"typedef struct %s\s*\{(.*?)\}\s*%s\s*;" % (name, name), hdr, re.S).group(1)
    out = []
    for decl in body.split(";"):
        decl = " ".join(decl.split())
        if not decl:
            continue
        ctype, rest = decl.split(" ", 1)
        for item in rest.split(","):
            item = item.strip()
            m = re.match(r"(\w+)\[(\d+)\]$", item)
            out.append((m.group(1), "NTuple{%s,%s}" % (m.group(2), CTYPE[ctype])) if m else (item, CTYPE[ctype]))
    return out


def julia_struct(name):
    src = open(os.path.join(ROOT, "julia", "RLSTDPB200.jl")).read()
    body = re.search(r"(?:mutable )?struct %s\n(.*?)\n(?:    %s\(\)|end)" % (name, name), src, re.S).group(1)
    return re.findall(r"(\w+)::([\w{},]+)", body)


def test_struct_mirrors_match_the_header():
    for c, j in (("snn_config", "SnnConfig"), ("snn_actor_config", "SnnActorConfig"), ("snn_da_event", "SnnDaEvent"),
                 ("snn_force_event", "SnnForceEvent"), ("snn_step_stats", "SnnStepStats")):
        assert julia_struct(j) == c_struct(c), (c, julia_struct(j), c_struct(c))


def test_every_ccall_symbol_is_declared():
    src = open(os.path.join(ROOT, "julia", "RLSTDPB200.jl")).read()
    hdr = _strip_comments(open(os.path.join(ROOT, "include", "snn_b200.h")).read())
    declared = set(re.findall(r"\b(snn_[a-z0-9_]+)\s*\(", hdr))
    called = set(re.findall(r"ccall\(\(:(\w+), LIB\)", src))
    assert called and called <= declared, called - declared
    assert {"snn_create", "snn_step", "snn_actor_episodes_env", "snn_partition_connect"} <= called


def _split_top(s):
    out, depth, cur = [], 0, ""
    for ch in s:
        if ch in "{(":
            depth += 1
        elif ch in "})":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip())
            cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_ccall_signatures_have_the_prototypes_arity_and_pointer_shape():
    src = open(os.path.join(ROOT, "julia", "RLSTDPB200.jl")).read()
    hdr = " ".join(_strip_comments(open(os.path.join(ROOT, "include", "snn_b200.h")).read()).split())
    protos = {m.group(1): _split_top(m.group(2)) for m in re.finditer(r"\b(snn_[a-z0-9_]+)\s*\(([^()]*)\)\s*;", hdr)}
    n = 0
    for m in re.finditer(r"ccall\(\(:(\w+), LIB\),\s*(\w+),\s*\(([^()]*)\)", src):
        name, ret, types = m.group(1), m.group(2), _split_top(m.group(3))
        cargs = [a for a in protos[name] if a != "void"]
        assert len(types) == len(cargs), (name, types, cargs)
        for jt, ca in zip(types, cargs):
            is_ptr_c = "*" in ca or ca.startswith("snn_exchange_fn")   # a function-pointer typedef
            is_ptr_j = jt.startswith(("Ptr{", "Ref{")) or jt == "Cstring"
            assert is_ptr_c == is_ptr_j, (name, jt, ca)
            if not is_ptr_c:
                assert CTYPE[ca.split()[0]] == jt, (name, jt, ca)
        assert ret in ("Int32", "Cstring", "Int64"), (name, ret)
        n += 1
    assert n >= 15
