#!/usr/bin/env python3
"""Writes simcu_embed.inc: the device-side headers as raw string literals for
the NVRTC translation unit (their #include lines are resolved by concatenation)."""
import sys

out = []
for name, path in (("kSrcProgram", "simcu_program.h"), ("kSrcCalc", "simcu_calc.h"),
                   ("kSrcDevice", "simcu_device.cuh"), ("kSrcJitKernel", "simcu_jit_kernel.cuh")):
    text = "".join(ln for ln in open(path) if not ln.lstrip().startswith("#include"))
    assert ')SIMCUSRC"' not in text
    # MSVC-free toolchain, but keep literals below the 64 KiB some compilers cap at
    chunks = [text[i:i + 12000] for i in range(0, len(text), 12000)]
    body = "\n".join('R"SIMCUSRC(%s)SIMCUSRC"' % c for c in chunks)
    out.append("static const char %s[] =\n%s;\n" % (name, body))
open(sys.argv[1], "w").write("\n".join(out))
