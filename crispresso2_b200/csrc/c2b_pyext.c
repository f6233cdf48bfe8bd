/* c2b_pyext.c -- CPython helpers for the host side of process_fastq: bulk construction of the variantCache.
 *
 * replaces: the per-unique-read Python work at the end of CRISPRessoCORE.process_fastq (:1956-1985) -- one dict per unique
 * read -- which at GPU alignment rates WAS the run time (r01: 35 us of Python per unique read around a 25 ns kernel).  The
 * engine's results stay in compact arrays; this module only creates, per aligned unique read, the key string and one lazy
 * dict object (crispresso2_b200/lazy.py: LazyVariant) that materialises the reference's dict / ResultsSlotsDict on first
 * access.
 *
 * Build: gcc -O2 -shared -fPIC -I<python include> -o crispresso2_b200/_c2b_pyext<EXT_SUFFIX> crispresso2_b200/csrc/c2b_pyext.c
 */
#define PY_SSIZE_T_CLEAN
#include <Python.h>
#include <pthread.h>
#include <stdint.h>
#include <stdlib.h>
#include <unistd.h>

/* str hashes of the keys, computed ahead by plain threads.  A dict insertion hashes its key; for a million 250-character keys
 * that is ~0.25 GB through SipHash on the thread that holds the GIL (measured: about half of the variantCache fill).  CPython's
 * str hash is _Py_HashBytes over the character data -- a pure function of the bytes and the process's hash secret -- so it is
 * computed here from the packed reads without the GIL and stored in the new str objects' hash field (where unicode_hash itself
 * caches it). */
typedef struct { const char *b; const int64_t *o; Py_hash_t *h; Py_ssize_t lo, hi; } hash_job;
static void *hash_worker(void *arg)
{
    hash_job *j = (hash_job *)arg;
    for (Py_ssize_t k = j->lo; k < j->hi; k++) j->h[k] = _Py_HashBytes(j->b + j->o[k], (Py_ssize_t)(j->o[k + 1] - j->o[k]));
    return NULL;
}
static Py_hash_t *hash_ahead(const char *b, const int64_t *o, Py_ssize_t n)
{
    if (n < 4096) return NULL;
    Py_hash_t *h = (Py_hash_t *)malloc((size_t)n * sizeof(Py_hash_t));
    if (!h) return NULL;
    long nc = sysconf(_SC_NPROCESSORS_ONLN);
    int T = nc > 32 ? 16 : nc > 3 ? (int)(nc / 2) : 1;
    pthread_t th[16];
    hash_job job[16];
    int started = 0;
    Py_BEGIN_ALLOW_THREADS
    for (int t = 0; t < T; t++) {
        job[t].b = b; job[t].o = o; job[t].h = h; job[t].lo = n * t / T; job[t].hi = n * (t + 1) / T;
        if (t == T - 1 || pthread_create(&th[t], NULL, hash_worker, &job[t]) != 0) { hash_worker(&job[t]); th[t] = 0; }
        else started |= 1 << t;
    }
    for (int t = 0; t < T; t++) if (started & (1 << t)) pthread_join(th[t], NULL);
    Py_END_ALLOW_THREADS
    return h;
}

/* make_keys(buf, off) -> list[str]: unique reads as the strings text-mode reading yields (UTF-8, surrogateescape) */
static PyObject *make_keys(PyObject *self, PyObject *args)
{
    Py_buffer buf, off;
    if (!PyArg_ParseTuple(args, "y*y*", &buf, &off)) return NULL;
    const Py_ssize_t n = off.len / 8 - 1;
    const int64_t *o = (const int64_t *)off.buf;
    const char *b = (const char *)buf.buf;
    PyObject *out = NULL;
    Py_hash_t *hashes = NULL;
    if (n < 0 || (n >= 0 && o[n < 0 ? 0 : n] > buf.len)) { PyErr_SetString(PyExc_ValueError, "offsets exceed the buffer"); goto done; }
    out = PyList_New(n);
    if (!out) goto done;
    hashes = hash_ahead(b, o, n);
    for (Py_ssize_t k = 0; k < n; k++) {
        const Py_ssize_t len = (Py_ssize_t)(o[k + 1] - o[k]);
        const unsigned char *p = (const unsigned char *)b + o[k];
        unsigned char hi = 0;
        for (Py_ssize_t x = 0; x < len; x++) hi |= p[x];
        PyObject *s;
        if (hi < 128) {                                    /* ASCII (every read that reaches the engine): no decoder */
            s = PyUnicode_New(len, 127);
            if (s) {
                memcpy(PyUnicode_1BYTE_DATA(s), p, (size_t)len);
                if (hashes) ((PyASCIIObject *)s)->hash = hashes[k];          /* the bytes ARE the character data */
            }
        } else s = PyUnicode_DecodeUTF8((const char *)p, len, "surrogateescape");
        if (!s) { Py_CLEAR(out); goto done; }
        PyList_SET_ITEM(out, k, s);
    }
done:
    free(hashes);
    PyBuffer_Release(&buf); PyBuffer_Release(&off);
    return out;
}

/* fill_cache(cache, keys, sel, counts, cls, value) -> number inserted
 * For every k with sel[k] == value: cache[keys[k]] = obj, obj = cls() with obj._k = k (the object answers ['count'] from
 * the batch's count array on first use, lazy.py).  Insertion order = k order (first-seen order of the unique reads).
 * The cyclic GC is paused meanwhile: half a million new containers would otherwise trigger full collections of a heap
 * that holds nothing collectable. */
static PyObject *fill_cache(PyObject *self, PyObject *args)
{
    PyObject *cache, *keys, *cls;
    Py_buffer sel, counts;
    int value;
    if (!PyArg_ParseTuple(args, "O!O!y*y*Oi", &PyDict_Type, &cache, &PyList_Type, &keys, &sel, &counts, &cls, &value)) return NULL;
    const Py_ssize_t n = PyList_GET_SIZE(keys);
    PyObject *res = NULL, *s_k = NULL, *staged = NULL, *target = cache;
    Py_ssize_t done = 0;
    const int gc_was = PyGC_Disable();
    if (sel.len < n || counts.len < 4 * n) { PyErr_SetString(PyExc_ValueError, "sel / counts shorter than keys"); goto out; }
    s_k = PyUnicode_InternFromString("_k");
    if (!s_k) goto out;
    {
        /* `cls` is a dict subclass with a `_k` slot (lazy.LazyVariant): instances are made with dict's tp_new directly (no
         * __init__ dispatch), `_k` is set through the slot's descriptor looked up once, and the new object is taken off the
         * cyclic GC's lists -- it holds an int until somebody materialises it (dict re-tracks itself when a container value
         * arrives), so a million of them must not turn every later collection into a 60 ms walk. */
        PyTypeObject *tp = (PyTypeObject *)cls;
        PyObject *descr = PyType_Check(cls) ? _PyType_Lookup(tp, s_k) : NULL;      /* borrowed */
        descrsetfunc setk = descr ? Py_TYPE(descr)->tp_descr_set : NULL;
        PyObject *empty = PyTuple_New(0);
        const int fast = PyType_Check(cls) && PyType_IsSubtype(tp, &PyDict_Type) && setk != NULL && empty != NULL;
        const uint8_t *m = (const uint8_t *)sel.buf;
        /* an empty plain dict is filled through a right-sized temporary: no rehash of half a million entries every time the table
         * grows, and dict.update() of an empty dict from such a table clones it wholesale (dict_merge's fast path) */
        Py_ssize_t want = 0;
        for (Py_ssize_t k = 0; k < n; k++) want += (m[k] == (uint8_t)value);
        if (PyDict_CheckExact(cache) && PyDict_GET_SIZE(cache) == 0 && want >= 4096) {
            staged = _PyDict_NewPresized(want);
            if (!staged) { Py_XDECREF(empty); goto out; }
            target = staged;
        }
        for (Py_ssize_t k = 0; k < n; k++) {
            if (m[k] != (uint8_t)value) continue;
            PyObject *o = fast ? PyDict_Type.tp_new(tp, empty, NULL) : PyObject_CallNoArgs(cls);
            if (!o) { Py_XDECREF(empty); goto out; }
            PyObject *ik = PyLong_FromSsize_t(k);
            int bad = !ik || (fast ? setk(descr, o, ik) < 0 : PyObject_SetAttr(o, s_k, ik) < 0);
            if (!bad && fast && PyObject_GC_IsTracked(o)) PyObject_GC_UnTrack(o);
            bad = bad || PyDict_SetItem(target, PyList_GET_ITEM(keys, k), o) < 0;
            Py_XDECREF(ik); Py_DECREF(o);
            if (bad) { Py_XDECREF(empty); goto out; }
            done++;
        }
        Py_XDECREF(empty);
        if (staged && PyDict_Update(cache, staged) < 0) goto out;
    }
    res = PyLong_FromSsize_t(done);
out:
    Py_XDECREF(s_k);
    Py_XDECREF(staged);
    if (gc_was) PyGC_Enable();
    PyBuffer_Release(&sel); PyBuffer_Release(&counts);
    return res;
}

/* slices(buf, starts, lens) -> list[str]: text of buf[starts[k] : starts[k] + lens[k]] (ASCII / latin-1) -- the aligned strings
 * of the allele table as Python objects, one C call for the whole column */
static PyObject *slices(PyObject *self, PyObject *args)
{
    Py_buffer buf, st, ln;
    if (!PyArg_ParseTuple(args, "y*y*y*", &buf, &st, &ln)) return NULL;
    const Py_ssize_t n = st.len / 8;
    const int64_t *o = (const int64_t *)st.buf;
    const int32_t *l = (const int32_t *)ln.buf;
    const char *b = (const char *)buf.buf;
    PyObject *out = NULL;
    if (ln.len / 4 < n) { PyErr_SetString(PyExc_ValueError, "lens shorter than starts"); goto done; }
    out = PyList_New(n);
    if (!out) goto done;
    for (Py_ssize_t k = 0; k < n; k++) {
        if (o[k] < 0 || l[k] < 0 || o[k] + l[k] > buf.len) { PyErr_SetString(PyExc_ValueError, "slice outside the buffer"); Py_CLEAR(out); goto done; }
        PyObject *s = PyUnicode_DecodeLatin1(b + o[k], (Py_ssize_t)l[k], NULL);
        if (!s) { Py_CLEAR(out); goto done; }
        PyList_SET_ITEM(out, k, s);
    }
done:
    PyBuffer_Release(&buf); PyBuffer_Release(&st); PyBuffer_Release(&ln);
    return out;
}

static PyMethodDef methods[] = {
    {"slices", slices, METH_VARARGS, "slices(buf, starts, lens) -> list of str"},
    {"make_keys", make_keys, METH_VARARGS, "make_keys(buf, off) -> list of str"},
    {"fill_cache", fill_cache, METH_VARARGS, "fill_cache(cache, keys, sel, counts, cls, value) -> int"},
    {NULL, NULL, 0, NULL}};

static struct PyModuleDef moddef = {PyModuleDef_HEAD_INIT, "_c2b_pyext", "bulk variantCache construction", -1, methods};

PyMODINIT_FUNC PyInit__c2b_pyext(void) { return PyModule_Create(&moddef); }
