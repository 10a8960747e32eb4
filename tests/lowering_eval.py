"""Evaluate a lowered KernelIR with NumPy (test helper: checks algorithm.py + cudagen.lower_statements
on the CPU, without a GPU, against the oracle's literal C kernels)."""
import numpy as np
import sympy as sp
from sympy.printing.numpy import NumPyPrinter

from pylbm_b200.cudagen import lower_statements


def evaluate(ir, inputs, scalars=None, cse=True):
    """inputs: list of arrays (one per input symbol); returns list of output arrays."""
    temps, outs = lower_statements(ir.statements, ir.outputs, cse=cse)
    pr = NumPyPrinter({"fully_qualified_modules": False})
    env = {"numpy": np, "sqrt": np.sqrt}
    for s, a in zip(ir.in_syms, inputs):
        env[str(s)] = a
    for k, v in (scalars or {}).items():
        env[k] = v
    lines = []
    for lhs, rhs in temps:
        lines.append("%s = %s" % (lhs, pr.doprint(rhs)))
    for i, o in enumerate(outs):
        lines.append("out_%d = %s" % (i, pr.doprint(o)))
    code = "\n".join(lines)
    exec(compile(code, "<lowered>", "exec"), env)
    shape = np.broadcast(*inputs).shape
    return [np.broadcast_to(np.asarray(env["out_%d" % i], dtype=float), shape) for i in range(len(outs))]
