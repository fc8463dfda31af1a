/* _hostpack: CPython helper of the drop-in modules/ layer (host glue, no arithmetic of the hot path).
 *
 * The reference's interface hands reads over as Python tuples of str (modules/cluster.py:207,
 * NGSpeciesID:58: (id, batch, accession, sequence, quality, score)); libngsid.so takes concatenated
 * byte arrays + offsets (ngsid_upload_reads). Doing that conversion with "".join(...).encode() costs
 * 125 ms per 100 k reads on the GPU box -- four times the clustering pass itself. pack_fields walks the
 * tuples once and copies the two str fields straight into two bytes objects.
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <stdint.h>
#include <string.h>

/* pack_fields(records, ia, ib) -> (bytes a, bytes b, bytes offsets)
 *   records: sequence of tuples / lists; fields ia and ib must be ASCII str of equal length per record.
 *   offsets: (n + 1) little-endian int64, offsets[i] = start of record i in both byte strings. */
static PyObject *pack_fields(PyObject *self, PyObject *args)
{
    PyObject *records, *fast = NULL, *ba = NULL, *bb = NULL, *bo = NULL, *res = NULL;
    Py_ssize_t ia, ib;
    if (!PyArg_ParseTuple(args, "Onn", &records, &ia, &ib)) return NULL;
    fast = PySequence_Fast(records, "records must be a sequence");
    if (!fast) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject **items = PySequence_Fast_ITEMS(fast);
    bo = PyBytes_FromStringAndSize(NULL, (n + 1) * (Py_ssize_t)sizeof(int64_t));
    if (!bo) goto done;
    int64_t *off = (int64_t *)PyBytes_AS_STRING(bo);
    int64_t total = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *fi = PySequence_Fast(items[i], "a record must be a tuple or list");
        if (!fi) goto done;
        const Py_ssize_t m = PySequence_Fast_GET_SIZE(fi);
        if (ia >= m || ib >= m || ia < 0 || ib < 0) { Py_DECREF(fi); PyErr_SetString(PyExc_IndexError, "field index out of range"); goto done; }
        PyObject *a = PySequence_Fast_ITEMS(fi)[ia], *b = PySequence_Fast_ITEMS(fi)[ib];
        if (!PyUnicode_Check(a) || !PyUnicode_Check(b)) { Py_DECREF(fi); PyErr_SetString(PyExc_TypeError, "fields must be str"); goto done; }
        if (!PyUnicode_IS_ASCII(a) || !PyUnicode_IS_ASCII(b)) { Py_DECREF(fi); PyErr_SetString(PyExc_ValueError, "sequence / quality strings must be ASCII"); goto done; }
        const Py_ssize_t la = PyUnicode_GET_LENGTH(a), lb = PyUnicode_GET_LENGTH(b);
        Py_DECREF(fi);
        if (la != lb) { PyErr_SetString(PyExc_ValueError, "sequence and quality of a record differ in length"); goto done; }
        off[i] = total;
        total += la;
    }
    off[n] = total;
    ba = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)total);
    bb = PyBytes_FromStringAndSize(NULL, (Py_ssize_t)total);
    if (!ba || !bb) goto done;
    {
        char *pa = PyBytes_AS_STRING(ba), *pb = PyBytes_AS_STRING(bb);
        for (Py_ssize_t i = 0; i < n; ++i) {
            PyObject *fi = PySequence_Fast(items[i], "a record must be a tuple or list");
            if (!fi) goto done;
            PyObject *a = PySequence_Fast_ITEMS(fi)[ia], *b = PySequence_Fast_ITEMS(fi)[ib];
            const Py_ssize_t la = (Py_ssize_t)(off[i + 1] - off[i]);
            /* ASCII str: one byte per character, PyUnicode_1BYTE_DATA is the text itself */
            memcpy(pa + off[i], PyUnicode_1BYTE_DATA(a), (size_t)la);
            memcpy(pb + off[i], PyUnicode_1BYTE_DATA(b), (size_t)la);
            Py_DECREF(fi);
        }
    }
    res = PyTuple_Pack(3, ba, bb, bo);
done:
    Py_XDECREF(fast); Py_XDECREF(ba); Py_XDECREF(bb); Py_XDECREF(bo);
    return res;
}


/* measure_fields(records, ia) -> total length of field ia over all records (ASCII str) */
static PyObject *measure_fields(PyObject *self, PyObject *args)
{
    PyObject *records;
    Py_ssize_t ia;
    if (!PyArg_ParseTuple(args, "On", &records, &ia)) return NULL;
    PyObject *fast = PySequence_Fast(records, "records must be a sequence");
    if (!fast) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject **items = PySequence_Fast_ITEMS(fast);
    long long total = 0;
    for (Py_ssize_t i = 0; i < n; ++i) {
        PyObject *fi = PySequence_Fast(items[i], "a record must be a tuple or list");
        if (!fi) { Py_DECREF(fast); return NULL; }
        if (ia < 0 || ia >= PySequence_Fast_GET_SIZE(fi) || !PyUnicode_Check(PySequence_Fast_ITEMS(fi)[ia])) {
            Py_DECREF(fi); Py_DECREF(fast);
            PyErr_SetString(PyExc_TypeError, "field must exist and be str");
            return NULL;
        }
        total += (long long)PyUnicode_GET_LENGTH(PySequence_Fast_ITEMS(fi)[ia]);
        Py_DECREF(fi);
    }
    Py_DECREF(fast);
    return PyLong_FromLongLong(total);
}

/* pack_fields_into(records, ia, ib, addr_a, addr_b, capacity) -> bytes offsets
 *   like pack_fields, but the two fields are copied to caller-owned memory (e.g. page-locked buffers from
 *   ngsid_pinned_alloc) at the integer addresses addr_a / addr_b, each with room for `capacity` bytes. */
static PyObject *pack_fields_into(PyObject *self, PyObject *args)
{
    PyObject *records, *fast = NULL, *bo = NULL, *res = NULL;
    Py_ssize_t ia, ib;
    unsigned long long addr_a, addr_b;
    long long cap;
    if (!PyArg_ParseTuple(args, "OnnKKL", &records, &ia, &ib, &addr_a, &addr_b, &cap)) return NULL;
    fast = PySequence_Fast(records, "records must be a sequence");
    if (!fast) return NULL;
    const Py_ssize_t n = PySequence_Fast_GET_SIZE(fast);
    PyObject **items = PySequence_Fast_ITEMS(fast);
    bo = PyBytes_FromStringAndSize(NULL, (n + 1) * (Py_ssize_t)sizeof(int64_t));
    if (!bo) goto done;
    {
        int64_t *off = (int64_t *)PyBytes_AS_STRING(bo);
        char *pa = (char *)(uintptr_t)addr_a, *pb = (char *)(uintptr_t)addr_b;
        int64_t total = 0;
        for (Py_ssize_t i = 0; i < n; ++i) {
            PyObject *fi = PySequence_Fast(items[i], "a record must be a tuple or list");
            if (!fi) goto done;
            const Py_ssize_t m = PySequence_Fast_GET_SIZE(fi);
            if (ia >= m || ib >= m || ia < 0 || ib < 0) { Py_DECREF(fi); PyErr_SetString(PyExc_IndexError, "field index out of range"); goto done; }
            PyObject *a = PySequence_Fast_ITEMS(fi)[ia], *b = PySequence_Fast_ITEMS(fi)[ib];
            if (!PyUnicode_Check(a) || !PyUnicode_Check(b)) { Py_DECREF(fi); PyErr_SetString(PyExc_TypeError, "fields must be str"); goto done; }
            if (!PyUnicode_IS_ASCII(a) || !PyUnicode_IS_ASCII(b)) { Py_DECREF(fi); PyErr_SetString(PyExc_ValueError, "sequence / quality strings must be ASCII"); goto done; }
            const Py_ssize_t la = PyUnicode_GET_LENGTH(a), lb = PyUnicode_GET_LENGTH(b);
            if (la != lb) { Py_DECREF(fi); PyErr_SetString(PyExc_ValueError, "sequence and quality of a record differ in length"); goto done; }
            if (total + la > cap) { Py_DECREF(fi); PyErr_SetString(PyExc_BufferError, "destination buffers too small"); goto done; }
            memcpy(pa + total, PyUnicode_1BYTE_DATA(a), (size_t)la);
            memcpy(pb + total, PyUnicode_1BYTE_DATA(b), (size_t)la);
            Py_DECREF(fi);
            off[i] = total;
            total += la;
        }
        off[n] = total;
    }
    res = bo; bo = NULL;
done:
    Py_XDECREF(fast); Py_XDECREF(bo);
    return res;
}

static PyMethodDef methods[] = {
    {"pack_fields", pack_fields, METH_VARARGS, "pack_fields(records, ia, ib) -> (bytes, bytes, offsets as int64 bytes)"},
    {"measure_fields", measure_fields, METH_VARARGS, "measure_fields(records, ia) -> total length of field ia"},
    {"pack_fields_into", pack_fields_into, METH_VARARGS, "pack_fields_into(records, ia, ib, addr_a, addr_b, capacity) -> offsets as int64 bytes"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_hostpack", "host-side packing helpers of ngspeciesid_b200", -1, methods};

PyMODINIT_FUNC PyInit__hostpack(void) { return PyModule_Create(&moddef); }
