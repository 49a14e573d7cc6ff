/* Build shim, force-included when compiling the reference's pyrSGM wrapper.
 * The wrapper calls PyArray_DATA on PyObject* (thirdparty/stereo-vision/reconstruction/
 * base/rSGM/pyrSGM.cpp:48 and friends), which numpy 2.x rejects at compile time.
 * This is test infrastructure only; it changes no behaviour. */
#include <Python.h>
#include <numpy/arrayobject.h>
#define PyArray_DATA(x) PyArray_DATA((PyArrayObject*)(x))
