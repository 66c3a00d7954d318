"""A small C# interpreter: TEST INFRASTRUCTURE ONLY (same rule as the rest of oracle/).

Why it exists: the reference (Storm-Tarran/LPR_381_Group_V22) is a .NET Framework 4.7.2 C# project and this
image has no .NET toolchain, so the reference cannot be built or run here.  `oracle/lpr_oracle.cpp` is a hand
restatement of its loops; a hand restatement can misread the source.  This package removes the reading step: it
parses the reference's OWN, UNMODIFIED `.cs` files (from /root/reference, in this container only) and executes
them -- C# grammar and operator precedence, int / double typing, IEEE-754 doubles (Python floats are the same
binary64 the CLR uses on x64), List<T> / LINQ / Math / string-formatting semantics of the Base Class Library
restated in `csrun.py` with the BCL rule each one follows.  `tests/golden/make_reference_run.py` uses it to run the
reference's solver classes on the reference's own fixtures and on seeded random models and commits what they
returned (`tests/golden/reference_run.json`); the oracle and the CUDA path are then compared with THAT.

Nothing under lpr_381_group_v22_b200/ imports this package (tests/test_cabi_symbols.py checks), and nothing on the
GPU box needs /root/reference: only the committed JSON travels.

    csparse.py  lexer + recursive-descent parser for the C# 7.3 subset the reference uses -> AST (tuples)
    csrun.py    tree-walking evaluator + the BCL subset (List, Dictionary, Stack, HashSet, LINQ, Math, String,
                StringBuilder, Console incl. SetOut, File, Path, Array, Tuple, Nullable, SafeHandle, Marshal, Encoding,
                exceptions, number formatting)
    pinvoke.py  `[DllImport] static extern` -> ctypes with the CLR's default marshalling: how the C# shims under
                csharp/ are executed against liblprb200.so (tests/test_csharp_shims*.py)

Also run with it: the reference's whole application (`Program.Main`, scripted keyboard) and the conformance tests of the
interpreter itself (tests/test_csharp_interpreter.py).
"""
from .csparse import parse_source  # noqa: F401
from .csrun import Interpreter, CsException  # noqa: F401
