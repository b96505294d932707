/* py_hotcalls.c — CPython binding of the C-ABI calls that sit on per-call hot paths of the Python view: evaluating (ct_do_buffer,
 * cc_buffer_release), building a node (ct_unary, ct_binary, ct_release) and reading back (ct_flat_array, ct_flat_buffer).
 * compute/scala_b200/cuda.py is a ctypes view of libcompute_cuda.so; ctypes spends ~0.55 us marshalling each call (more with pointer
 * arguments), which is a third of a 3.2 us step of BASELINE config 1 and most of a 20 us `flatArray` of a small tensor (both host-bound,
 * DESIGN.md section 8). These go through METH_O / METH_FASTCALL functions instead (~0.1 us).
 * No logic lives here: arguments in, status or handle out, the GIL released around calls that reach the driver, as ctypes does. If
 * this module is not built, cuda.py binds the same entry points through ctypes (a binding choice, not a compute path: both end in
 * libcompute_cuda.so). */
#define PY_SSIZE_T_CLEAN
#include <Python.h>

#include "../../../include/compute_cuda.h"

/* do_buffer(tensor_handle) -> buffer handle (> 0), or the negative cc_status of the failed call */
static PyObject* hot_do_buffer(PyObject* self, PyObject* arg) {
  (void)self;
  const unsigned long long t = PyLong_AsUnsignedLongLong(arg);
  if (t == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
  cc_buffer out = 0;
  int st;
  Py_BEGIN_ALLOW_THREADS
  st = ct_do_buffer((ct_tensor)t, &out, NULL);
  Py_END_ALLOW_THREADS
  if (st != CC_OK) return PyLong_FromLong(st < 0 ? st : -st);
  return PyLong_FromUnsignedLongLong((unsigned long long)out);
}

/* buffer_release(buffer_handle) -> cc_status */
static PyObject* hot_buffer_release(PyObject* self, PyObject* arg) {
  (void)self;
  const unsigned long long b = PyLong_AsUnsignedLongLong(arg);
  if (b == (unsigned long long)-1 && PyErr_Occurred()) return NULL;
  int st;
  Py_BEGIN_ALLOW_THREADS
  st = cc_buffer_release((cc_buffer)b);
  Py_END_ALLOW_THREADS
  return PyLong_FromLong(st);
}

static int as_u64(PyObject* o, unsigned long long* out) {
  *out = PyLong_AsUnsignedLongLong(o);
  return !(*out == (unsigned long long)-1 && PyErr_Occurred());
}
static PyObject* handle_or_status(int st, unsigned long long h) {
  if (st != CC_OK) return PyLong_FromLong(st < 0 ? st : -st);
  return PyLong_FromUnsignedLongLong(h);
}

/* unary(op, tensor) / binary(op, lhs, rhs) -> tensor handle, or the negative cc_status (graph construction: no device work) */
static PyObject* hot_unary(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
  (void)self;
  unsigned long long t;
  if (nargs != 2) return PyErr_Format(PyExc_TypeError, "unary(op, tensor)");
  const long op = PyLong_AsLong(args[0]);
  if ((op == -1 && PyErr_Occurred()) || !as_u64(args[1], &t)) return NULL;
  ct_tensor out = 0;
  const int st = ct_unary((int)op, (ct_tensor)t, &out);
  return handle_or_status(st, (unsigned long long)out);
}
static PyObject* hot_binary(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
  (void)self;
  unsigned long long l, r;
  if (nargs != 3) return PyErr_Format(PyExc_TypeError, "binary(op, lhs, rhs)");
  const long op = PyLong_AsLong(args[0]);
  if ((op == -1 && PyErr_Occurred()) || !as_u64(args[1], &l) || !as_u64(args[2], &r)) return NULL;
  ct_tensor out = 0;
  const int st = ct_binary((int)op, (ct_tensor)l, (ct_tensor)r, &out);
  return handle_or_status(st, (unsigned long long)out);
}
/* tensor_release(tensor) -> cc_status (may free device buffers of a whole sub-graph: GIL released) */
static PyObject* hot_tensor_release(PyObject* self, PyObject* arg) {
  (void)self;
  unsigned long long t;
  if (!as_u64(arg, &t)) return NULL;
  int st;
  Py_BEGIN_ALLOW_THREADS
  st = ct_release((ct_tensor)t);
  Py_END_ALLOW_THREADS
  return PyLong_FromLong(st);
}
/* flat_array_into(tensor, writable float32 buffer) -> cc_status: evaluates and reads back into the caller's memory (T:1111-1118) */
static PyObject* hot_flat_array_into(PyObject* self, PyObject* const* args, Py_ssize_t nargs) {
  (void)self;
  unsigned long long t;
  if (nargs != 2) return PyErr_Format(PyExc_TypeError, "flat_array_into(tensor, buffer)");
  if (!as_u64(args[0], &t)) return NULL;
  Py_buffer view;
  if (PyObject_GetBuffer(args[1], &view, PyBUF_WRITABLE | PyBUF_C_CONTIGUOUS) != 0) return NULL;
  int st;
  Py_BEGIN_ALLOW_THREADS
  st = ct_flat_array((ct_tensor)t, (float*)view.buf, (uint64_t)view.len / 4);
  Py_END_ALLOW_THREADS
  PyBuffer_Release(&view);
  return PyLong_FromLong(st);
}
/* flat_buffer(tensor) -> (host pointer, writable memoryview over the pinned block) or the negative cc_status (T:1099-1109).
 * The block stays valid until flat_buffer_release(pointer); the memoryview must not be used after that. */
static PyObject* hot_flat_buffer(PyObject* self, PyObject* arg) {
  (void)self;
  unsigned long long t;
  if (!as_u64(arg, &t)) return NULL;
  float* host = NULL;
  uint64_t n = 0;
  int st;
  Py_BEGIN_ALLOW_THREADS
  st = ct_flat_buffer((ct_tensor)t, &host, &n);
  Py_END_ALLOW_THREADS
  if (st != CC_OK) return PyLong_FromLong(st < 0 ? st : -st);
  static char empty[4];
  PyObject* mv = PyMemoryView_FromMemory(host && n ? (char*)host : empty, (Py_ssize_t)(host && n ? n * 4 : 0), PyBUF_WRITE);
  if (!mv) {
    ct_flat_buffer_release(host);
    return NULL;
  }
  return Py_BuildValue("(KN)", (unsigned long long)(uintptr_t)host, mv);
}
static PyObject* hot_flat_buffer_release(PyObject* self, PyObject* arg) {
  (void)self;
  unsigned long long p;
  if (!as_u64(arg, &p)) return NULL;
  return PyLong_FromLong(ct_flat_buffer_release((float*)(uintptr_t)p));
}

static PyMethodDef methods[] = {
    {"unary", (PyCFunction)(void (*)(void))hot_unary, METH_FASTCALL, "ct_unary(op, tensor) -> tensor handle, or a negative cc_status"},
    {"binary", (PyCFunction)(void (*)(void))hot_binary, METH_FASTCALL, "ct_binary(op, lhs, rhs) -> tensor handle, or a negative cc_status"},
    {"tensor_release", hot_tensor_release, METH_O, "ct_release(tensor) -> cc_status"},
    {"flat_array_into", (PyCFunction)(void (*)(void))hot_flat_array_into, METH_FASTCALL, "ct_flat_array(tensor, writable buffer) -> cc_status"},
    {"flat_buffer", hot_flat_buffer, METH_O, "ct_flat_buffer(tensor) -> (pointer, memoryview), or a negative cc_status"},
    {"flat_buffer_release", hot_flat_buffer_release, METH_O, "ct_flat_buffer_release(pointer) -> cc_status"},
    {"do_buffer", hot_do_buffer, METH_O, "ct_do_buffer(tensor) -> buffer handle, or a negative cc_status"},
    {"buffer_release", hot_buffer_release, METH_O, "cc_buffer_release(buffer) -> cc_status"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_hotcalls", "hot-path bindings of libcompute_cuda.so", -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__hotcalls(void) { return PyModule_Create(&module); }
