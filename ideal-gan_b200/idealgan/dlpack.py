"""DLPack capsule inspection with ctypes: the zero-copy handshake between a tensor framework and the C ABI.

Any object with `__dlpack__` (torch, TensorFlow through tf.experimental.dlpack.to_dlpack, CuPy, JAX) hands over a
PyCapsule named "dltensor" that wraps a DLManagedTensor.  `tensor_info` reads the device pointer, shape, strides,
dtype and device out of it without copying or consuming it, which is all libidealgan needs (raw pointers + sizes).
Struct layout: DLPack v0.x, https://github.com/dmlc/dlpack (dlpack.h), as shipped by torch 2.x and TF 2.x.
"""
import ctypes as C
from collections import namedtuple

kDLCPU, kDLCUDA, kDLCUDAHost, kDLCUDAManaged = 1, 2, 3, 13
kDLInt, kDLUInt, kDLFloat, kDLBfloat, kDLComplex = 0, 1, 2, 4, 5


class DLDevice(C.Structure):
    _fields_ = [("device_type", C.c_int), ("device_id", C.c_int)]


class DLDataType(C.Structure):
    _fields_ = [("code", C.c_uint8), ("bits", C.c_uint8), ("lanes", C.c_uint16)]


class DLTensor(C.Structure):
    _fields_ = [("data", C.c_void_p), ("device", DLDevice), ("ndim", C.c_int), ("dtype", DLDataType),
                ("shape", C.POINTER(C.c_int64)), ("strides", C.POINTER(C.c_int64)), ("byte_offset", C.c_uint64)]


class DLManagedTensor(C.Structure):
    pass


DLManagedTensor._fields_ = [("dl_tensor", DLTensor), ("manager_ctx", C.c_void_p),
                            ("deleter", C.CFUNCTYPE(None, C.POINTER(DLManagedTensor)))]

TensorInfo = namedtuple("TensorInfo", "ptr shape strides dtype_code bits device_type device_id contiguous")

_PyCapsule_GetPointer = C.pythonapi.PyCapsule_GetPointer
_PyCapsule_GetPointer.restype = C.c_void_p
_PyCapsule_GetPointer.argtypes = [C.py_object, C.c_char_p]
_PyCapsule_IsValid = C.pythonapi.PyCapsule_IsValid
_PyCapsule_IsValid.restype = C.c_int
_PyCapsule_IsValid.argtypes = [C.py_object, C.c_char_p]


def capsule_info(capsule):
    """Read a *live* "dltensor" capsule.  The capsule is not consumed: the caller keeps ownership."""
    if not _PyCapsule_IsValid(capsule, b"dltensor"):
        raise ValueError("not an unconsumed DLPack capsule (expected name 'dltensor')")
    mt = C.cast(_PyCapsule_GetPointer(capsule, b"dltensor"), C.POINTER(DLManagedTensor)).contents
    t = mt.dl_tensor
    shape = tuple(int(t.shape[i]) for i in range(t.ndim))
    if t.strides:
        strides = tuple(int(t.strides[i]) for i in range(t.ndim))
    else:
        strides, acc = [], 1
        for n in reversed(shape):
            strides.append(acc)
            acc *= n
        strides = tuple(reversed(strides))
    expect, acc = [], 1
    for n in reversed(shape):
        expect.append(acc)
        acc *= n
    contiguous = all(n == 1 or s == e for n, s, e in zip(shape, strides, reversed(expect)))
    return TensorInfo((t.data or 0) + t.byte_offset, shape, strides, t.dtype.code, t.dtype.bits, t.device.device_type,
                      t.device.device_id, contiguous)


def tensor_info(obj):
    """TensorInfo of any DLPack exporter; the exporting tensor must outlive the use of the returned pointer."""
    if hasattr(obj, "__dlpack__"):
        cap = obj.__dlpack__()
    else:                                            # TensorFlow eager tensors
        import tensorflow as tf
        cap = tf.experimental.dlpack.to_dlpack(obj)
    info = capsule_info(cap)
    del cap                                          # unconsumed capsule: its destructor runs the producer's deleter
    return info


def require_f32_cuda(info, name):
    if info.dtype_code != kDLFloat or info.bits != 32:
        raise ValueError(f"{name}: dtype must be float32")
    if info.device_type not in (kDLCUDA, kDLCUDAManaged):
        raise ValueError(f"{name}: tensor must live on a CUDA device (this path has no CPU fallback)")
    if not info.contiguous:
        raise ValueError(f"{name}: tensor must be C-contiguous")
    return info
