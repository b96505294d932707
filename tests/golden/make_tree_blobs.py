"""Writes tests/golden/tree_blobs/<case>.bin: the tree blob the C++ mirror (csrc/tensor.cpp) hands to cc_compile_ex for each case of
tree_blob_cases.py, with parameter ids normalised to first-emission ordinals (the mirror uses tensor addresses). Needs no GPU:
    python tests/golden/make_tree_blobs.py
tests/test_scala_twin.py holds the mirror, the Python twin of CudaTreeWriter.scala and (on a machine with a JVM)
scala/.../CudaTreeWriterSpec.scala to these files."""
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def mirror_blob(cuda, what, kind) -> bytes:
    T = cuda.Tensor
    if kind == "tensor":
        return what.treeBlob()
    if kind[0] == "join":
        return (T.join(what) if kind[1] is None else T.join(what, kind[1])).treeBlob()
    if kind[0] == "reduce":
        return what.reduce(kind[1]).treeBlob()
    raise ValueError(kind)


def main() -> None:
    from golden.tree_blob_cases import cases
    from scala_twin import normalise_ids

    from compute.scala_b200 import cuda

    out_dir = os.path.join(HERE, "tree_blobs")
    os.makedirs(out_dir, exist_ok=True)
    for name, (what, kind) in cases(cuda.Tensor).items():
        with open(os.path.join(out_dir, name + ".bin"), "wb") as f:
            f.write(normalise_ids(mirror_blob(cuda, what, kind)))
    print(len(os.listdir(out_dir)), "blobs in", out_dir)


if __name__ == "__main__":
    main()
