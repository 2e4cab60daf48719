"""
Golden Quil text of the reference's classical decoder (css_code.quil_classical_correct / _detect,
css_code.py:649-713), produced by running the UNMODIFIED reference under a *recording* pyquil stub:
``Program += instruction`` collects instructions, ``gates.MOVE/AND/XOR/IOR/NOT`` render the text pyquil
would print (``MOVE scratch[0] codeword[1]``).  Test infrastructure only; run here once, the fixture is
committed because /root/reference does not exist on the GPU box.

    python oracle/gen_quil_golden.py        # writes tests/golden/quil_classical_golden.json
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
            def __init__(self):
                self.instructions = []
            def __iadd__(self, other):
                if isinstance(other, str):
                    self.instructions.append(other)
                else:
                    self.instructions.extend(other)
                return self
        def get_qc(*a, **k):
            raise RuntimeError('stub')
    """,
    "gates.py": """
        def _two(op):
            return lambda a, b: f"{op} {a} {b}"
        MOVE, AND, XOR, IOR = _two("MOVE"), _two("AND"), _two("XOR"), _two("IOR")
        def NOT(a):
            return f"NOT {a}"
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
            pass
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
    from pyquil.quilatom import MemoryReference
    from quil_classical import MemoryChunk

    sys.path.insert(0, REPO)
    from quantum_css_codes_b200 import codes

    out = {}
    for name in ("steane", "qrm15"):
        h1, h2 = [np.array(h) for h in getattr(codes, name)()]
        code = ref.CSSCode(h1, h2)
        for tag, h, table in (("c1", code.parity_check_c1, code._c1_syndromes),
                              ("c2", code.parity_check_c2, code._c2_syndromes)):
            if len(table) > 64:
                continue                                     # keep the fixture small (QRM c2 has 576 entries)
            m, n = h.shape
            codeword = MemoryChunk(MemoryReference("codeword", declared_size=n), 0, n)
            errors = MemoryChunk(MemoryReference("errors", declared_size=n), 0, n)
            scratch = MemoryChunk(MemoryReference("scratch", declared_size=m + 2), 0, m + 2)
            prog = Program()
            ref.quil_classical_correct(prog, codeword, errors, scratch, h, table)
            out[f"{name}_{tag}_correct"] = list(prog.instructions)
            prog = Program()
            ref.quil_classical_detect(prog, codeword, errors, MemoryReference("outcome")[0], scratch, h)
            out[f"{name}_{tag}_detect"] = list(prog.instructions)
    dst = os.path.join(REPO, "tests", "golden", "quil_classical_golden.json")
    with open(dst, "w") as fh:
        json.dump(out, fh)
    print("wrote", dst, {k: len(v) for k, v in out.items()})


if __name__ == "__main__":
    main()
