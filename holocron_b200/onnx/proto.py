"""The subset of the ONNX protobuf schema (onnx/onnx.proto, IR version 7) an exported inference graph needs, declared
programmatically so that files can be written (and read back) with the stock ``google.protobuf`` runtime - the ``onnx`` python
package is not a dependency. Field numbers and enum values are those of the published schema:

    ModelProto{ir_version=1, producer_name=2, producer_version=3, domain=4, model_version=5, doc_string=6, graph=7,
               opset_import=8}            OperatorSetIdProto{domain=1, version=2}
    GraphProto{node=1, name=2, initializer=5, doc_string=10, input=11, output=12, value_info=13}
    NodeProto{input=1, output=2, name=3, op_type=4, attribute=5, doc_string=6, domain=7}
    AttributeProto{name=1, f=2, i=3, s=4, t=5, floats=7, ints=8, strings=9, type=20}
    TensorProto{dims=1, data_type=2, float_data=4, int32_data=5, int64_data=7, name=8, raw_data=9}
    ValueInfoProto{name=1, type=2}  TypeProto{tensor_type=1{elem_type=1, shape=2}}  TensorShapeProto{dim=1{dim_value=1, dim_param=2}}
"""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_F = descriptor_pb2.FieldDescriptorProto
_T = {"int64": _F.TYPE_INT64, "int32": _F.TYPE_INT32, "float": _F.TYPE_FLOAT, "string": _F.TYPE_STRING, "bytes": _F.TYPE_BYTES}

# AttributeProto.AttributeType / TensorProto.DataType values of the schema
ATTR_FLOAT, ATTR_INT, ATTR_STRING, ATTR_TENSOR, ATTR_FLOATS, ATTR_INTS, ATTR_STRINGS = 1, 2, 3, 4, 6, 7, 8
DT_FLOAT, DT_INT32, DT_INT64, DT_BOOL = 1, 6, 7, 9

_PKG = "hb_onnx"     # private package name: never clashes with a real `onnx` installation's descriptor pool entries


def _msg(fd, name, fields, nested=None):
    """fields: (name, number, type | '.pkg.Message', repeated)."""
    m = fd.message_type.add() if nested is None else nested.nested_type.add()
    m.name = name
    for fname, number, ftype, repeated in fields:
        f = m.field.add()
        f.name, f.number = fname, number
        f.label = _F.LABEL_REPEATED if repeated else _F.LABEL_OPTIONAL
        if ftype in _T:
            f.type = _T[ftype]
        else:
            f.type, f.type_name = _F.TYPE_MESSAGE, f".{_PKG}.{ftype}"
    return m


def _build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "hb_onnx_subset.proto", _PKG, "proto2"
    _msg(fd, "OperatorSetIdProto", [("domain", 1, "string", False), ("version", 2, "int64", False)])
    _msg(fd, "TensorProto", [("dims", 1, "int64", True), ("data_type", 2, "int32", False), ("float_data", 4, "float", True),
                             ("int32_data", 5, "int32", True), ("int64_data", 7, "int64", True), ("name", 8, "string", False),
                             ("raw_data", 9, "bytes", False)])
    shape = _msg(fd, "TensorShapeProto", [("dim", 1, "TensorShapeProto.Dimension", True)])
    _msg(fd, "Dimension", [("dim_value", 1, "int64", False), ("dim_param", 2, "string", False)], nested=shape)
    tp = _msg(fd, "TypeProto", [("tensor_type", 1, "TypeProto.Tensor", False)])
    _msg(fd, "Tensor", [("elem_type", 1, "int32", False), ("shape", 2, "TensorShapeProto", False)], nested=tp)
    _msg(fd, "ValueInfoProto", [("name", 1, "string", False), ("type", 2, "TypeProto", False)])
    _msg(fd, "AttributeProto", [("name", 1, "string", False), ("f", 2, "float", False), ("i", 3, "int64", False),
                                ("s", 4, "bytes", False), ("t", 5, "TensorProto", False), ("floats", 7, "float", True),
                                ("ints", 8, "int64", True), ("strings", 9, "bytes", True), ("type", 20, "int32", False)])
    _msg(fd, "NodeProto", [("input", 1, "string", True), ("output", 2, "string", True), ("name", 3, "string", False),
                           ("op_type", 4, "string", False), ("attribute", 5, "AttributeProto", True),
                           ("doc_string", 6, "string", False), ("domain", 7, "string", False)])
    _msg(fd, "GraphProto", [("node", 1, "NodeProto", True), ("name", 2, "string", False), ("initializer", 5, "TensorProto", True),
                            ("doc_string", 10, "string", False), ("input", 11, "ValueInfoProto", True),
                            ("output", 12, "ValueInfoProto", True), ("value_info", 13, "ValueInfoProto", True)])
    _msg(fd, "ModelProto", [("ir_version", 1, "int64", False), ("producer_name", 2, "string", False),
                            ("producer_version", 3, "string", False), ("domain", 4, "string", False),
                            ("model_version", 5, "int64", False), ("doc_string", 6, "string", False),
                            ("graph", 7, "GraphProto", False), ("opset_import", 8, "OperatorSetIdProto", True)])
    pool = descriptor_pool.DescriptorPool()
    pool.Add(fd)
    return {n: message_factory.GetMessageClass(pool.FindMessageTypeByName(f"{_PKG}.{n}"))
            for n in ("ModelProto", "GraphProto", "NodeProto", "AttributeProto", "TensorProto", "ValueInfoProto", "TypeProto",
                      "TensorShapeProto", "OperatorSetIdProto")}


_CLASSES = _build()
ModelProto = _CLASSES["ModelProto"]
GraphProto = _CLASSES["GraphProto"]
NodeProto = _CLASSES["NodeProto"]
AttributeProto = _CLASSES["AttributeProto"]
TensorProto = _CLASSES["TensorProto"]
ValueInfoProto = _CLASSES["ValueInfoProto"]
