"""Platform invoke for the interpreter: `[DllImport] static extern` methods of an interpreted C# class call a real shared
library through ctypes, with the default marshalling of the CLR for the types the shims under csharp/ use.

TEST INFRASTRUCTURE ONLY (see oracle/csharp/__init__.py).  This is how the P/Invoke shims of csharp/ -- the drop-in a
maintainer adds to the reference, never compiled here because the image has no .NET -- are EXECUTED against
liblprb200.so: tests/test_csharp_shims.py.

Marshalling rules restated (ECMA-335 II.15.5, "Default Marshaling Behavior" of the .NET Framework):
    int / long / double / byte        by value (int32 / int64 / double / uint8)
    IntPtr                            void*; IntPtr.Zero is NULL
    string                            const char* (CharSet.Ansi: the platform encoding, UTF-8 here), null -> NULL
    T[] and T[,] of blittable T       pinned, passed as a pointer to the first element (row-major for T[,]); the callee's
                                      writes are visible to the caller with or without [Out]; null -> NULL
    string[]                          array of const char*
    out T / ref T                     pointer to a T; out starts zero-initialised
"""
import ctypes as C

from .csrun import CsArray, Ref

_SCALAR = {"int": C.c_int, "long": C.c_int64, "double": C.c_double, "byte": C.c_ubyte, "bool": C.c_int,
           "uint": C.c_uint, "ulong": C.c_uint64, "short": C.c_short, "float": C.c_float, "IntPtr": C.c_void_p}


class NativeLibrary:
    def __init__(self, cdll):
        self.lib = cdll
        self.calls = []          # names of the entry points called, in order (the tests assert on these)

    def call(self, name, params, values, rettype):
        fn = getattr(self.lib, name)
        cargs, after, keep = [], [], []
        for (ty, pname, _default, mod), v in zip(params, values):
            base = ty[1]
            if ty[3]:                                   # array
                if mod in ("out", "ref"):
                    raise NotImplementedError(f"{name}: out/ref arrays are not used by the shims")
                if v is None:
                    cargs.append(None)
                    continue
                if type(v) is not CsArray:
                    raise TypeError(f"{name}.{pname}: expected an array")
                if base == "string":
                    enc = [None if s is None else s.encode("utf-8") for s in v.data]
                    arr = (C.c_char_p * max(1, len(enc)))(*enc)
                    keep.append(enc)
                    cargs.append(arr)
                    continue
                ct = _SCALAR[base]
                arr = (ct * max(1, len(v.data)))(*v.data)
                cargs.append(arr)
                after.append((v, arr, base))
                continue
            if mod in ("out", "ref"):
                if type(v) is not Ref:
                    raise TypeError(f"{name}.{pname}: expected an lvalue")
                ct = _SCALAR[base]
                cell = ct()
                if mod == "ref":
                    cur = v.get()
                    cell.value = cur if base != "IntPtr" else (cur or None)
                cargs.append(C.byref(cell))
                after.append((v, cell, base))
                continue
            if base == "string":
                cargs.append(None if v is None else v.encode("utf-8"))
            elif base == "IntPtr":
                cargs.append(C.c_void_p(v or None))
            elif base in _SCALAR:
                cargs.append(_SCALAR[base](int(v) if base not in ("double", "float") else float(v)))
            else:
                raise NotImplementedError(f"{name}.{pname}: no marshalling for {base}")
        rbase = rettype[1] if rettype is not None else "void"
        fn.restype = None if rbase == "void" else _SCALAR[rbase]
        fn.argtypes = None
        self.calls.append(name)
        r = fn(*cargs)
        for target, cval, base in after:
            if type(target) is Ref:
                x = cval.value
                target.set(0 if (base == "IntPtr" and x is None) else x)
            else:
                if base in ("double", "float"):
                    target.data[:] = [float(x) for x in cval[:len(target.data)]]
                else:
                    target.data[:] = [int(x) for x in cval[:len(target.data)]]
        if rbase == "IntPtr":
            return r or 0
        return r
