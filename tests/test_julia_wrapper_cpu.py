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
