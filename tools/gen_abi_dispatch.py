#!/usr/bin/env python
"""Regenerate cudanavierstokes_b200/csrc/abi_dispatch.cpp and cudns_abi.h from the prototypes of include/cudns.h: the device-side
entry points exist once per precision (cudns64_* / cudns32_*), the public symbol forwards to the copy the handle belongs to."""
import os, re
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CS = os.path.join(ROOT, "cudanavierstokes_b200", "csrc")
# every prototype of the header that takes (or creates) a solver handle
hdr = re.sub(r"/\*.*?\*/", "", open(os.path.join(ROOT, "include", "cudns.h")).read(), flags=re.S)
protos = [(m.group(1), " ".join(m.group(2).split())) for m in re.finditer(r"\bint\s+cudns_(\w+)\s*\(([^;]*?)\)\s*;", hdr, flags=re.S)
          if "cudns_handle" in m.group(2) and not m.group(1).startswith("team_")]      # (cudns_team_* works on top of the public symbols)
out = open(os.path.join(CS, "abi_dispatch.cpp")).read().split('extern "C" {')[0] + 'extern "C" {\n'
for n, params in protos:
    args = [re.findall(r"([A-Za-z_][A-Za-z0-9_]*)\s*(?:\[[^\]]*\])?$", a.strip())[0] for a in params.split(",")]
    sel = "p && p->precision == 1" if n == "create" else "%s && *(const int *)%s == 1" % (args[0], args[0])
    out += "int cudns64_%s(%s);\nint cudns32_%s(%s);\n" % (n, params, n, params)
    out += "int cudns_%s(%s) { return (%s) ? cudns32_%s(%s) : cudns64_%s(%s); }\n" % (n, params, sel, n, ", ".join(args), n, ", ".join(args))
out += '\n}  // extern "C"\n'
open(os.path.join(CS, "abi_dispatch.cpp"), "w").write(out)
abi = open(os.path.join(CS, "cudns_abi.h")).read().split("#define cudns_solver")[0] + "#define cudns_solver CUDNS_PREC_NAME(solver)\n"
abi += "".join("#define cudns_%s CUDNS_PREC_NAME(%s)\n" % (n, n) for n, _ in protos)
open(os.path.join(CS, "cudns_abi.h"), "w").write(abi)
print(len(protos), "entry points")
