"""
TEST INFRASTRUCTURE (oracle side) -- not part of the product path.

Decoder for the reference's ``.sym`` expression files
(``/root/reference/src/dynamical_systems/{energy,gradient}_functions/*.sym``).

Those files are dill pickles of SymPy-lambdified functions whose code objects
were created by CPython 3.8.  They unpickle on 3.12 but cannot execute
(SURVEY.md section 0, fact 4).  The bytecode is straight line and uses only
11 opcodes, so we replay it symbolically and emit equivalent Python *source*.
Nothing from the reference is copied into this repository: decoding happens at
oracle run time, reading the files where they lie.

Loader call sites this replaces (``dl_load`` = ``dill.load``):
``lorenz_96.py:399,407,415``, ``lorenz_63.py:204-220``,
``double_well.py:111-127``, ``ornstein_uhlenbeck.py:105-121``.
"""
import io
import pickle

# CPython 3.8 opcode numbers used by the 18 files (wordcode: 2 bytes/instr).
_LOAD_FAST, _STORE_FAST, _LOAD_CONST = 0x7C, 0x7D, 0x64
_UNARY_NEGATIVE, _RETURN_VALUE, _BUILD_LIST = 0x0B, 0x53, 0x67
_BINARY = {0x17: "+", 0x18: "-", 0x14: "*", 0x1B: "/", 0x13: "**"}


class _CodeSpec(object):
    """Holds the 16 positional fields dill passes to ``_create_code``."""

    def __init__(self, *fields):
        if len(fields) != 16:
            raise ValueError(f"unexpected code layout: {len(fields)} fields")
        (self.argcount, self.posonly, self.kwonly, self.nlocals, self.stacksize,
         self.flags, self.code, self.consts, self.names, self.varnames,
         self.filename, self.name, self.firstlineno, self.lnotab,
         self.freevars, self.cellvars) = fields


def _stub_create_function(code, fglobals, name=None, defaults=None,
                          closure=None, fdict=None, kwdefaults=None):
    return code


class _StubUnpickler(pickle.Unpickler):
    """Never builds a real code object: dill's factories are replaced by stubs."""

    def find_class(self, module, name):
        if module == "dill._dill" and name == "_create_code":
            return _CodeSpec
        if module == "dill._dill" and name == "_create_function":
            return _stub_create_function
        raise pickle.UnpicklingError(f"unexpected global {module}.{name}")


def load_code_spec(fileobj) -> _CodeSpec:
    spec = _StubUnpickler(io.BytesIO(fileobj.read())).load()
    if not isinstance(spec, _CodeSpec):
        raise TypeError("pickle did not contain a function")
    if spec.names or spec.freevars or spec.cellvars:
        raise ValueError("expression uses globals or closures")
    return spec


def decode_to_source(spec: _CodeSpec, func_name: str = "_expr") -> str:
    """Symbolically execute the 3.8 wordcode and return ``def`` source."""
    args = spec.varnames[:spec.argcount]
    lines = [f"def {func_name}({', '.join(args)}):"]
    stack = []
    code = spec.code
    for pc in range(0, len(code), 2):
        op, arg = code[pc], code[pc + 1]
        if op == _LOAD_FAST:
            stack.append(spec.varnames[arg])
        elif op == _LOAD_CONST:
            stack.append(repr(spec.consts[arg]))
        elif op == _STORE_FAST:
            lines.append(f"    {spec.varnames[arg]} = {stack.pop()}")
        elif op == _UNARY_NEGATIVE:
            stack.append(f"(-{stack.pop()})")
        elif op in _BINARY:
            rhs = stack.pop()
            lhs = stack.pop()
            stack.append(f"({lhs} {_BINARY[op]} {rhs})")
        elif op == _BUILD_LIST:
            items = stack[len(stack) - arg:]
            del stack[len(stack) - arg:]
            stack.append("[" + ", ".join(items) + "]")
        elif op == _RETURN_VALUE:
            lines.append(f"    return {stack.pop()}")
        else:
            raise ValueError(f"unsupported opcode 0x{op:02x} at pc={pc}")
    if stack:
        raise ValueError("stack not empty at end of code")
    return "\n".join(lines) + "\n"


def load(fileobj):
    """Drop-in for ``dill.load`` on a ``.sym`` file: returns a Python function."""
    spec = load_code_spec(fileobj)
    src = decode_to_source(spec)
    namespace = {}
    exec(compile(src, f"<decoded {getattr(fileobj, 'name', 'sym')}>", "exec"),
         namespace)
    return namespace["_expr"]
