"""TEST INFRASTRUCTURE — a line-by-line Python transliteration of scala/cuda/.../CudaTreeWriter.scala (and of the three places
CudaTensors.scala calls it from: enqueueClosure, join, reduce), walking the oracle's tree classes, which restate Trees.scala's case classes
(oracle/reference.py: FloatLiteral R:373-380, ArrayParameter R:755-823, Transform R:676-690, Extract R:660-672, Concatenate R:953-973,
unary R:384-470, binary R:472-618).

No JVM exists in the build environment, so the Scala writer cannot be run; this twin pins what it must write: tests/test_scala_twin.py
compares the twin's blobs with the blobs of the C++ mirror (csrc/tensor.cpp, ct_tree_blob), byte for byte once parameter ids are
normalised, checks that the library's structural cache treats both as the same kernel, and checks them against the committed golden
blobs under tests/golden/tree_blobs/ (which scala/.../CudaTreeWriterSpec.scala reads on a machine that does have a JVM).
Keep the two files in step: every method below names its Scala counterpart."""
from __future__ import annotations

import struct

from oracle import reference as ref

UNARY = {"Exp": 10, "Log": 11, "Abs": 12, "Tanh": 13, "Sqrt": 14, "UnaryMinus": 15}
BINARY = {"Min": 20, "Max": 21, "Plus": 22, "Minus": 23, "Times": 24, "Div": 25, "Percent": 26}
MONOID = {"+": 22, "min": 20, "max": 21, "*": 24}


class CudaTreeWriter:
    def __init__(self):
        self.body = bytearray()
        self.number_of_nodes = 0
        self.node_of_tree: dict[int, int] = {}  # IdentityHashMap[Tree, Integer]
        self.definition_field_offset: dict[int, int] = {}
        self.parameters: list = []  # the producing tensors in order of first emission; id in the blob = index + 1
        self.parameter_node: dict[int, int] = {}
        self._keep = []  # (keeps id()s unique for the lifetime of the writer)

    # private def begin(kind, payloadBytes)
    def _begin(self, kind: int) -> int:
        self.body += struct.pack("<I", kind)
        self.number_of_nodes += 1
        return self.number_of_nodes - 1

    # private def operands(tree)
    @staticmethod
    def _operands(tree):
        if isinstance(tree, ref.Unary):
            return [tree.operand0]
        if isinstance(tree, ref.Binary):
            return [tree.operand0, tree.operand1]
        if isinstance(tree, (ref.Extract, ref.Transform)):
            return [tree.array]
        if isinstance(tree, ref.Concatenate):
            return list(tree.elements)
        if isinstance(tree, (ref.FloatLiteral, ref.ArrayParameter)):
            return []
        raise ValueError(f"{type(tree).__name__} cannot be reached from a Tensor")

    # private def emit(tree)
    def _emit(self, tree) -> int:
        if isinstance(tree, ref.FloatLiteral):
            node = self._begin(1)
            self.body += struct.pack("<f", float(tree.value))
            return node
        if isinstance(tree, ref.ArrayParameter):
            node = self._begin(2)
            self.parameters.append(tree.id)
            self.parameter_node[id(tree.id)] = node
            self.body += struct.pack("<Q", len(self.parameters))
            self.body += struct.pack("<f", float(tree.padding.value))
            self.body += struct.pack("<I", len(tree.shape))
            for s in tree.shape:
                self.body += struct.pack("<i", s)
            self.definition_field_offset[node] = len(self.body)
            self.body += struct.pack("<i", -1)
            return node
        if isinstance(tree, ref.Transform):
            array_node = self.node_of_tree[id(tree.array)]
            rows = len(tree.array.shape)  # rankOfArray: Transform sits directly over an ArrayParameter
            assert rows > 0 and len(tree.matrix) % rows == 0
            node = self._begin(3)
            self.body += struct.pack("<III", array_node, rows, len(tree.matrix) // rows)
            for v in tree.matrix:
                self.body += struct.pack("<d", float(v))
            return node
        if isinstance(tree, ref.Extract):
            node = self._begin(4)
            self.body += struct.pack("<I", self.node_of_tree[id(tree.array)])
            return node
        if isinstance(tree, ref.Concatenate):
            node = self._begin(5)
            self.body += struct.pack("<I", len(tree.elements))
            for e in tree.elements:
                self.body += struct.pack("<I", self.node_of_tree[id(e)])
            return node
        if isinstance(tree, ref.Unary):
            node = self._begin(UNARY[tree.op])
            self.body += struct.pack("<I", self.node_of_tree[id(tree.operand0)])
            return node
        if isinstance(tree, ref.Binary):
            node = self._begin(BINARY[tree.op])
            self.body += struct.pack("<II", self.node_of_tree[id(tree.operand0)], self.node_of_tree[id(tree.operand1)])
            return node
        raise ValueError(type(tree).__name__)

    # def write(tree): post-order, operands left to right, memoised by identity, explicit stack
    def write(self, tree) -> int:
        if id(tree) in self.node_of_tree:
            return self.node_of_tree[id(tree)]
        stack = [(tree, False)]
        while stack:
            node, expanded = stack.pop()
            if id(node) in self.node_of_tree:
                continue
            pending = [o for o in self._operands(node) if id(o) not in self.node_of_tree]
            if expanded or not pending:
                self._keep.append(node)
                self.node_of_tree[id(node)] = self._emit(node)
            else:
                stack.append((node, True))
                for o in reversed(pending):
                    stack.append((o, False))
        return self.node_of_tree[id(tree)]

    # def concatenateAt(elements, position)
    def concatenate_at(self, elements, position: int) -> int:
        node = self._begin(6)
        self.body += struct.pack("<II", len(elements), position)
        for e in elements:
            self.body += struct.pack("<I", e)
        return node

    # def reduce(monoid, operand, operandShape)
    def reduce(self, monoid: int, operand: int, operand_shape) -> int:
        node = self._begin(30)
        self.body += struct.pack("<III", monoid, operand, len(operand_shape))
        for s in operand_shape:
            self.body += struct.pack("<i", s)
        return node

    # def attachDefinition(tensor, closure)
    def attach_definition(self, tensor, closure) -> None:
        node = self.parameter_node[id(tensor)]
        definition_root = self.write(closure)
        struct.pack_into("<i", self.body, self.definition_field_offset[node], definition_root)

    # def finish(root, outShape)
    def finish(self, root: int, out_shape) -> bytes:
        head = struct.pack("<IIII", 0x31544343, self.number_of_nodes, root, len(out_shape))
        for s in out_shape:
            head += struct.pack("<i", s)
        return head + bytes(self.body)


# ---- CudaTensors.scala: the three call sites ------------------------------------------------------------------------------------


def _attach_definitions(writer: CudaTreeWriter) -> None:
    """private def attachDefinitions(writer): every MAIN-tree parameter whose producing tensor is an InlineTensor carries its closure"""
    for tensor in list(writer.parameters):
        if isinstance(tensor, ref._Inline):
            writer.attach_definition(tensor, tensor.closure())


def blob_of(tensor, monoid: str | None = None, join_dimension: int | None = None) -> bytes:
    """the blob `compile(shape)(writeRoot)` hands to cc_compile_ex for a tensor of the oracle (= of Tensors.scala):
    InlineTensor.plan (closure), Tensor.join / join(…, dimension) (Concatenate / ConcatenateAt), Tensor.reduce over an inline operand"""
    w = CudaTreeWriter()
    if monoid is not None:  # Tensor.reduce: writer.reduce(monoid.kind, writer.write(closure.tree), shape), out shape []
        root = w.reduce(MONOID[monoid], w.write(tensor.closure()), tensor.shape)
        out_shape = ()
    elif isinstance(tensor, (list, tuple)):  # Tensor.join(tensors[, dimension])
        tensors = list(tensor)
        rank = len(tensors[0].shape)
        if join_dimension is not None and join_dimension != rank:
            root = w.concatenate_at([w.write(t.closure()) for t in tensors], join_dimension)
            out_shape = tuple(tensors[0].shape[:join_dimension]) + (len(tensors),) + tuple(tensors[0].shape[join_dimension:])
        else:
            # (the Scala side writes trees.tuple.join(closures).tree: elements first, then the Concatenate record)
            root = w.write(ref.Concatenate([t.closure() for t in tensors]))
            out_shape = tuple(tensors[0].shape) + (len(tensors),)
    else:
        root = w.write(tensor.closure())
        out_shape = tuple(tensor.shape)
    _attach_definitions(w)
    return w.finish(root, out_shape)


def normalise_ids(blob: bytes) -> bytes:
    """replaces every ArrayParameter id by its first-emission ordinal (the C++ mirror uses the tensor's address, the Scala writer 1 + ordinal)"""
    out = bytearray(blob)
    magic, n_nodes, _root, rank = struct.unpack_from("<IIII", blob, 0)
    assert magic == 0x31544343
    p = 16 + 4 * rank
    ordinal = 0
    for _ in range(n_nodes):
        (kind,) = struct.unpack_from("<I", blob, p)
        p += 4
        if kind == 1:
            p += 4
        elif kind == 2:
            ordinal += 1
            struct.pack_into("<Q", out, p, ordinal)
            (r,) = struct.unpack_from("<I", blob, p + 12)
            p += 8 + 4 + 4 + 4 * r + 4
        elif kind == 3:
            _a, rows, cols = struct.unpack_from("<III", blob, p)
            p += 12 + 8 * rows * cols
        elif kind == 4 or 10 <= kind <= 15:
            p += 4
        elif kind == 5:
            (n,) = struct.unpack_from("<I", blob, p)
            p += 4 + 4 * n
        elif kind == 6:
            (n,) = struct.unpack_from("<I", blob, p)
            p += 8 + 4 * n
        elif 20 <= kind <= 26:
            p += 8
        elif kind == 30:
            (r,) = struct.unpack_from("<I", blob, p + 8)
            p += 12 + 4 * r
        else:
            raise ValueError(kind)
    assert p == len(blob)
    return bytes(out)
