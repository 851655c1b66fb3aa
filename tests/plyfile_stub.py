"""Minimal stand-in for the third-party `plyfile` package (absent from this image) -- TEST INFRASTRUCTURE ONLY.

Implements exactly the subset the reference uses in scene/gaussian_model.py:285-313,364-418
(`PlyElement.describe(structured_array, name)`, `PlyData([el]).write(path)`, `PlyData.read(path)`,
`plydata.elements[0][prop]`, `plydata.elements[0].properties[i].name`) with plyfile's default on-disk format:
`format binary_little_endian 1.0`, one `property <type> <name>` line per structured-dtype field, rows as packed
little-endian records.  It lets the tests run the reference's UNMODIFIED save_ply / load_ply against
instascene_b200.io's files."""
import numpy as np

_TYPES = {"f4": "float", "f8": "double", "i1": "char", "u1": "uchar", "i2": "short", "u2": "ushort", "i4": "int", "u4": "uint"}
_REV = {v: k for k, v in _TYPES.items()}
_REV.update({"float32": "f4", "float64": "f8", "int8": "i1", "uint8": "u1", "int16": "i2", "uint16": "u2", "int32": "i4", "uint32": "u4"})


class PlyProperty:
    def __init__(self, name, val_dtype):
        self.name, self.val_dtype = name, val_dtype


class PlyElement:
    def __init__(self, name, data):
        self.name, self.data = name, data
        self.properties = tuple(PlyProperty(n, data.dtype[n].str[1:]) for n in data.dtype.names)

    @staticmethod
    def describe(data, name):
        if not isinstance(data, np.ndarray) or data.dtype.names is None:
            raise TypeError("only structured numpy arrays are supported")
        return PlyElement(name, data)

    def __getitem__(self, key):
        return self.data[key]

    def __len__(self):
        return len(self.data)


class PlyData:
    def __init__(self, elements=(), text=False, byte_order="="):
        self.elements = list(elements)
        if text:
            raise NotImplementedError("the reference writes binary files")

    def write(self, path):
        with open(path, "wb") as f:
            lines = ["ply", "format binary_little_endian 1.0"]
            for el in self.elements:
                lines.append(f"element {el.name} {len(el.data)}")
                lines += [f"property {_TYPES[p.val_dtype]} {p.name}" for p in el.properties]
            lines.append("end_header")
            f.write(("\n".join(lines) + "\n").encode("ascii"))
            for el in self.elements:
                le = el.data.astype(el.data.dtype.newbyteorder("<"), copy=False)
                f.write(np.ascontiguousarray(le).tobytes())

    @staticmethod
    def read(path):
        with open(path, "rb") as f:
            if f.readline().strip() != b"ply":
                raise ValueError("not a PLY file")
            fmt, elements, cur = None, [], None
            while True:
                line = f.readline()
                if not line:
                    raise ValueError("unexpected end of header")
                tok = line.decode("ascii").split()
                if not tok or tok[0] == "comment" or tok[0] == "obj_info":
                    continue
                if tok[0] == "format":
                    fmt = tok[1]
                elif tok[0] == "element":
                    cur = [tok[1], int(tok[2]), []]
                    elements.append(cur)
                elif tok[0] == "property":
                    if tok[1] == "list":
                        raise NotImplementedError("list properties are not used by the Gaussian point cloud")
                    cur[2].append((tok[2], _REV[tok[1]]))
                elif tok[0] == "end_header":
                    break
            out = []
            for name, count, props in elements:
                if fmt == "ascii":
                    rows = [f.readline().split() for _ in range(count)]
                    data = np.array([tuple(float(x) for x in r) for r in rows], dtype=[(n, "<" + t) for n, t in props])
                else:
                    bo = "<" if fmt == "binary_little_endian" else ">"
                    dt = np.dtype([(n, bo + t) for n, t in props])
                    data = np.frombuffer(f.read(count * dt.itemsize), dtype=dt, count=count).copy()
                out.append(PlyElement(name, data))
            return PlyData(out)
