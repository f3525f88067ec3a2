/* CSR <-> List[List[int]] for the list-returning BPE API (CPython C API, no numpy / torch headers).
 *
 * The reference's `BEASTBsplineBPETokenizer.encode` returns one Python list of ids per trajectory
 * (`beast/beast_bspline_bpe_tokenizer.py:175-198`, `encoding.ids` per row) and `reconstruct_traj` / `decode` take
 * that ragged list back (`:200-247`).  The kernels work on CSR (flat int32 ids + int64 offsets); at 65 536
 * trajectories x ~118 ids the conversion in Python (one `tolist()`, 65 536 slices, `np.fromiter` over a chain)
 * cost 200-300 ms against < 10 ms of GPU work.  Here it is two tight loops:
 *
 *   split_rows(flat, offsets)  -> list of lists; every distinct id is ONE shared int object (ints are immutable),
 *                                 so a row costs a PyList_New + one pointer store and incref per id
 *   flatten_rows(rows)         -> (flat bytearray of int32, offsets bytearray of int64 [n+1]) from a list / tuple of
 *                                 lists / tuples of Python ints; returns None when some row is of another type
 *                                 (the caller then takes the general path), raises ValueError for ids outside int32
 *
 * Host glue only: no tokenizer arithmetic happens here.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

#define ID_CACHE_MAX (1 << 20) /* ids above this get a fresh int object each */

static PyObject *split_rows(PyObject *self, PyObject *args) {
    Py_buffer flat, off;
    if (!PyArg_ParseTuple(args, "y*y*", &flat, &off)) return NULL;
    PyObject *result = NULL;
    PyObject **cache = NULL;
    int64_t cache_n = 0;
    const int32_t *ids = (const int32_t *)flat.buf;
    const int64_t *offs = (const int64_t *)off.buf;
    const int64_t n_ids = (int64_t)(flat.len / 4);
    const int64_t n_rows = (int64_t)(off.len / 8) - 1;
    if (flat.len % 4 || off.len % 8 || n_rows < 0) {
        PyErr_SetString(PyExc_ValueError, "split_rows: flat must be int32 and offsets int64 [n+1]");
        goto done;
    }
    if (offs[0] != 0 || offs[n_rows] != n_ids) {
        PyErr_SetString(PyExc_ValueError, "split_rows: offsets do not cover the id array");
        goto done;
    }
    int32_t hi = -1;
    for (int64_t i = 0; i < n_ids; ++i)
        if (ids[i] > hi) hi = ids[i];
    cache_n = hi < ID_CACHE_MAX ? (int64_t)hi + 1 : ID_CACHE_MAX;
    cache = (PyObject **)PyMem_Calloc((size_t)(cache_n > 0 ? cache_n : 1), sizeof(PyObject *));
    if (!cache) {
        PyErr_NoMemory();
        goto done;
    }
    result = PyList_New((Py_ssize_t)n_rows);
    if (!result) goto done;
    for (int64_t r = 0; r < n_rows; ++r) {
        const int64_t b = offs[r], e = offs[r + 1];
        if (e < b || e > n_ids) {
            PyErr_SetString(PyExc_ValueError, "split_rows: offsets must be non-decreasing");
            Py_CLEAR(result);
            goto done;
        }
        PyObject *row = PyList_New((Py_ssize_t)(e - b));
        if (!row) {
            Py_CLEAR(result);
            goto done;
        }
        PyList_SET_ITEM(result, (Py_ssize_t)r, row); /* result owns the row from here on */
        for (int64_t i = b; i < e; ++i) {
            const int32_t v = ids[i];
            PyObject *o;
            if (v >= 0 && v < cache_n) {
                o = cache[v];
                if (!o) {
                    o = PyLong_FromLong(v);
                    if (!o) {
                        Py_CLEAR(result);
                        goto done;
                    }
                    cache[v] = o; /* the cache keeps one reference until the end of the call */
                }
                Py_INCREF(o);
            } else {
                o = PyLong_FromLong(v);
                if (!o) {
                    Py_CLEAR(result);
                    goto done;
                }
            }
            PyList_SET_ITEM(row, (Py_ssize_t)(i - b), o);
        }
    }
done:
    if (cache) {
        for (int64_t i = 0; i < cache_n; ++i) Py_XDECREF(cache[i]);
        PyMem_Free(cache);
    }
    PyBuffer_Release(&flat);
    PyBuffer_Release(&off);
    return result;
}

static PyObject *flatten_rows(PyObject *self, PyObject *rows) {
    const int rows_is_list = PyList_CheckExact(rows);
    if (!rows_is_list && !PyTuple_CheckExact(rows)) Py_RETURN_NONE;
    const Py_ssize_t n = rows_is_list ? PyList_GET_SIZE(rows) : PyTuple_GET_SIZE(rows);
    int64_t total = 0;
    for (Py_ssize_t r = 0; r < n; ++r) {
        PyObject *row = rows_is_list ? PyList_GET_ITEM(rows, r) : PyTuple_GET_ITEM(rows, r);
        if (PyList_CheckExact(row))
            total += PyList_GET_SIZE(row);
        else if (PyTuple_CheckExact(row))
            total += PyTuple_GET_SIZE(row);
        else
            Py_RETURN_NONE;
    }
    PyObject *flat = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)(total * 4));
    PyObject *off = PyByteArray_FromStringAndSize(NULL, (Py_ssize_t)((n + 1) * 8));
    if (!flat || !off) {
        Py_XDECREF(flat);
        Py_XDECREF(off);
        return NULL;
    }
    int32_t *ids = (int32_t *)PyByteArray_AS_STRING(flat);
    int64_t *offs = (int64_t *)PyByteArray_AS_STRING(off);
    int64_t k = 0;
    offs[0] = 0;
    for (Py_ssize_t r = 0; r < n; ++r) {
        PyObject *row = rows_is_list ? PyList_GET_ITEM(rows, r) : PyTuple_GET_ITEM(rows, r);
        const int is_list = PyList_CheckExact(row);
        const Py_ssize_t m = is_list ? PyList_GET_SIZE(row) : PyTuple_GET_SIZE(row);
        for (Py_ssize_t i = 0; i < m; ++i) {
            PyObject *o = is_list ? PyList_GET_ITEM(row, i) : PyTuple_GET_ITEM(row, i);
            if (!PyLong_CheckExact(o)) { /* bools, numpy scalars, floats: the general path decides */
                Py_DECREF(flat);
                Py_DECREF(off);
                Py_RETURN_NONE;
            }
            int overflow = 0;
            const long v = PyLong_AsLongAndOverflow(o, &overflow);
            if (overflow || v < INT32_MIN || v > INT32_MAX) {
                Py_DECREF(flat);
                Py_DECREF(off);
                PyErr_SetString(PyExc_ValueError, "BPE token id out of range");
                return NULL;
            }
            ids[k++] = (int32_t)v;
        }
        offs[r + 1] = k;
    }
    PyObject *out = PyTuple_Pack(2, flat, off);
    Py_DECREF(flat);
    Py_DECREF(off);
    return out;
}

static PyMethodDef methods[] = {
    {"split_rows", split_rows, METH_VARARGS, "split_rows(flat int32 buffer, offsets int64 buffer) -> list of lists of int"},
    {"flatten_rows", flatten_rows, METH_O, "flatten_rows(rows) -> (flat int32 bytearray, offsets int64 bytearray) or None"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef module = {PyModuleDef_HEAD_INIT, "_pylists", "CSR <-> ragged Python lists", -1, methods};

PyMODINIT_FUNC PyInit__pylists(void) { return PyModule_Create(&module); }
