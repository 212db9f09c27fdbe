/* csrpack.c -- host-side helper of the drop-in layer: python list of id lists -> CSR arrays in one C loop.
 *
 * The reference's callers hand the path python lists (src/wmd.py:34-45: `xs1`, `xs2` are lists of tokenizer-id
 * lists built by src/loader.py:46-60).  Flattening them with numpy costs ~45 ns per token of interpreter work
 * (225 us for the two sides of one 256-pair Yelp batch -- more than the GPU call); this does it at ~5 ns per token.
 *
 *   pack(docs, width) -> (ids: bytes, off: bytes)     width = 4: int32 ids, 8: int64 ids; off is int64[len(docs) + 1]
 * ids that do not fit the width become -1 (out of vocabulary for the library).
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>

static PyObject *pack(PyObject *self, PyObject *args)
{
    PyObject *docs_in;
    int width = 4;
    if (!PyArg_ParseTuple(args, "O|i", &docs_in, &width)) return NULL;
    if (width != 4 && width != 8) { PyErr_SetString(PyExc_ValueError, "width must be 4 or 8"); return NULL; }
    PyObject *docs = PySequence_Fast(docs_in, "expected a sequence of id sequences");
    if (!docs) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(docs);
    PyObject *off_b = PyBytes_FromStringAndSize(NULL, (n + 1) * 8);
    if (!off_b) { Py_DECREF(docs); return NULL; }
    int64_t *off = (int64_t *)PyBytes_AS_STRING(off_b);
    int64_t total = 0;
    off[0] = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        const Py_ssize_t l = PySequence_Size(PySequence_Fast_GET_ITEM(docs, i));
        if (l < 0) { Py_DECREF(docs); Py_DECREF(off_b); return NULL; }
        total += l;
        off[i + 1] = total;
    }
    PyObject *ids_b = PyBytes_FromStringAndSize(NULL, total * width);
    if (!ids_b) { Py_DECREF(docs); Py_DECREF(off_b); return NULL; }
    char *out = PyBytes_AS_STRING(ids_b);
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *d = PySequence_Fast(PySequence_Fast_GET_ITEM(docs, i), "expected a sequence of ids");
        if (!d) goto fail;
        const Py_ssize_t l = PySequence_Fast_GET_SIZE(d);
        if (l != off[i + 1] - off[i]) { Py_DECREF(d); PyErr_SetString(PyExc_RuntimeError, "document changed size while packing"); goto fail; }
        PyObject **items = PySequence_Fast_ITEMS(d);
        for (Py_ssize_t k = 0; k < l; ++k) {
            int overflow = 0;
            long long v = PyLong_AsLongLongAndOverflow(items[k], &overflow);
            if (v == -1 && !overflow && PyErr_Occurred()) {          /* not an int: try __index__ (numpy scalars) */
                PyErr_Clear();
                PyObject *ix = PyNumber_Index(items[k]);
                if (!ix) { Py_DECREF(d); goto fail; }
                v = PyLong_AsLongLongAndOverflow(ix, &overflow);
                Py_DECREF(ix);
            }
            if (overflow) v = -1;
            if (width == 4) { if (v < INT32_MIN || v > INT32_MAX) v = -1; *(int32_t *)out = (int32_t)v; }
            else *(int64_t *)out = (int64_t)v;
            out += width;
        }
        Py_DECREF(d);
    }
    Py_DECREF(docs);
    return Py_BuildValue("(NN)", ids_b, off_b);
fail:
    Py_DECREF(docs); Py_DECREF(off_b); Py_DECREF(ids_b);
    return NULL;
}

static PyMethodDef methods[] = {
    { "pack", pack, METH_VARARGS, "pack(docs, width=4) -> (ids bytes, int64 offsets bytes)" },
    { NULL, NULL, 0, NULL }
};
static struct PyModuleDef moddef = { PyModuleDef_HEAD_INIT, "_csrpack", "python list of id lists -> CSR arrays", -1, methods };
PyMODINIT_FUNC PyInit__csrpack(void) { return PyModule_Create(&moddef); }
