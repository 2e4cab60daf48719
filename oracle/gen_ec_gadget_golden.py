"""
Golden instruction stream of the reference's Steane error-correction gadget (CSSCode.error_correct,
css_code.py:436-470), produced by running the UNMODIFIED reference under a *recording* pyquil stub: gates render
as the text pyquil would print (``CNOT d0 a0``, ``H a3``, ``MEASURE a3 scratch[3]``, ``XOR scratch[0] dx[0]``),
``Program.while_do`` / ``if_then`` are recorded as nested blocks.  This is what pins SURVEY 8 f-4: tests/
test_ec_gadget.py propagates injected Pauli errors through THIS instruction stream (CNOT / H / MEASURE on a
Pauli frame, the classical part through quil_text.run) and requires the outcome of oracle/ec_rounds.py's round
model, error for error.  Test infrastructure only; run here once, the fixture is committed because
/root/reference does not exist on the GPU box.

    python oracle/gen_ec_gadget_golden.py        # writes tests/golden/ec_gadget_golden.json
"""

import json
import os
import sys
import tempfile
import textwrap

import numpy as np

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REFERENCE = os.environ.get("QCSS_REFERENCE", "/root/reference")

STUB = {
    "__init__.py": """
        class Program:
            def __init__(self, *parts):
                self.instructions = []
                for part in parts:
                    self += part
            def __iadd__(self, other):
                if isinstance(other, Program):
                    self.instructions.extend(other.instructions)
                elif isinstance(other, (str, dict)):
                    self.instructions.append(other)
                else:
                    for item in other:
                        self += item
                return self
            def __add__(self, other):
                out = Program(self)
                out += other
                return out
            def if_then(self, cond, then_branch, else_branch=None):
                assert else_branch is None
                self.instructions.append({"if": str(cond), "then": Program(then_branch).instructions})
                return self
            def while_do(self, cond, body):
                self.instructions.append({"while": str(cond), "body": Program(body).instructions})
                return self
        def get_qc(*a, **k):
            raise RuntimeError('stub')
    """,
    "gates.py": """
        def _one(op):
            return lambda a: f"{op} {a}"
        def _two(op):
            return lambda a, b: f"{op} {a} {b}"
        MOVE, AND, XOR, IOR = _two("MOVE"), _two("AND"), _two("XOR"), _two("IOR")
        NOT = _one("NOT")
        I, X, Y, Z, H, S = _one("I"), _one("X"), _one("Y"), _one("Z"), _one("H"), _one("S")
        CNOT, CZ, MEASURE = _two("CNOT"), _two("CZ"), _two("MEASURE")
        QUANTUM_GATES = {"I": I, "X": X, "Y": Y, "Z": Z, "H": H, "S": S, "CNOT": CNOT, "CZ": CZ}
    """,
    "paulis.py": "class PauliTerm:\n    pass\ndef ID():\n    raise RuntimeError('stub')\nsX = sY = sZ = ID\n",
    "quil.py": "from pyquil import Program\n",
    "quilatom.py": """
        class MemoryReference:
            def __init__(self, name, offset=0, declared_size=None):
                self.name, self.offset, self.declared_size = name, offset, declared_size
            def __getitem__(self, index):
                return MemoryReference(self.name, self.offset + index)
            def __str__(self):
                return f"{self.name}[{self.offset}]"
        class Qubit:
            pass
        class QubitPlaceholder:
            def __init__(self, name=None):
                self.name = name
            def __str__(self):
                return self.name
    """,
    "quilbase.py": "class Gate:\n    pass\n",
}


def main():
    tmp = tempfile.mkdtemp()
    pkg = os.path.join(tmp, "pyquil")
    os.makedirs(pkg)
    for name, body in STUB.items():
        with open(os.path.join(pkg, name), "w") as fh:
            fh.write(textwrap.dedent(body))
    sys.path.insert(0, REFERENCE)
    sys.path.insert(0, tmp)
    import warnings
    warnings.simplefilter("ignore")
    import css_code as ref                                   # the reference module
    from pyquil import Program
    from pyquil.quilatom import MemoryReference, QubitPlaceholder
    from qecc import CodeBlock
    from quil_classical import MemoryChunk

    sys.path.insert(0, REPO)
    from quantum_css_codes_b200 import codes

    out = {}
    for name in ("steane", "shor9"):
        h1, h2 = [np.array(h) for h in getattr(codes, name)()]
        code = ref.CSSCode(h1, h2)
        n = code.n

        def block(tag):
            return CodeBlock([QubitPlaceholder(f"{tag}{i}") for i in range(n)],
                             MemoryChunk(MemoryReference(f"{tag}x", declared_size=n), 0, n),
                             MemoryChunk(MemoryReference(f"{tag}z", declared_size=n), 0, n))

        # (error_correct_scratch_size = 2 n - max(r_1, r_2) + 4 is too small for the reference's own _error_detect_x
        #  when r_1 != r_2, e.g. Shor-9; a larger buffer is accepted)
        size = max(code.error_correct_scratch_size, 2 * n + max(code.r_1, code.r_2) + 4)
        scratch = MemoryChunk(MemoryReference("scratch", declared_size=size), 0, size)
        prog = Program()
        code.error_correct(prog, block("d"), block("a"), block("b"), scratch)
        out[name] = {"n": n, "scratch": size, "program": prog.instructions}
    dst = os.path.join(REPO, "tests", "golden", "ec_gadget_golden.json")
    with open(dst, "w") as fh:
        json.dump(out, fh, separators=(",", ":"))

    def count(instr):
        return sum(1 + (count(i.get("body", i.get("then", []))) if isinstance(i, dict) else 0) for i in instr)
    print("wrote", dst, {k: count(v["program"]) for k, v in out.items()}, os.path.getsize(dst), "bytes")


if __name__ == "__main__":
    main()
