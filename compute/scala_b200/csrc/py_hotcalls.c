/* py_hotcalls.c — CPython binding of the two C-ABI calls a launch loop makes every step (ct_do_buffer, cc_buffer_release).
 * compute/scala_b200/cuda.py is a ctypes view of libcompute_cuda.so; ctypes spends ~0.55 us marshalling each call, which is a third of
 * a 3.3 us step of BASELINE config 1 (host-bound). These two go through a METH_O function instead (~0.1 us).
 * No logic lives here: arguments in, status or handle out, the GIL released around the call as ctypes does. If this module is not
 * built, cuda.py binds the same two entry points through ctypes (a binding choice, not a compute path: both end in libcompute_cuda.so). */
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

static PyMethodDef methods[] = {
    {"do_buffer", hot_do_buffer, METH_O, "ct_do_buffer(tensor) -> buffer handle, or a negative cc_status"},
    {"buffer_release", hot_buffer_release, METH_O, "cc_buffer_release(buffer) -> cc_status"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_hotcalls", "hot-path bindings of libcompute_cuda.so", -1, methods, NULL, NULL, NULL, NULL};

PyMODINIT_FUNC PyInit__hotcalls(void) { return PyModule_Create(&module); }
