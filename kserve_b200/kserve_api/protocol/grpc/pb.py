"""Message classes of the Open Inference Protocol gRPC schema (package `inference`), built at import time from a
programmatic FileDescriptorProto: there is no protoc / grpc_tools in this environment, and the schema is the public
standard the reference vendors as python/kserve/kserve/protocol/grpc/grpc_predict_v2.proto (field numbers from there)."""
from google.protobuf import descriptor_pb2, descriptor_pool, message_factory

_T = descriptor_pb2.FieldDescriptorProto
_SCALAR = {"bool": _T.TYPE_BOOL, "int32": _T.TYPE_INT32, "int64": _T.TYPE_INT64, "uint32": _T.TYPE_UINT32,
           "uint64": _T.TYPE_UINT64, "float": _T.TYPE_FLOAT, "double": _T.TYPE_DOUBLE, "string": _T.TYPE_STRING,
           "bytes": _T.TYPE_BYTES}


def _field(msg, name, number, typ, repeated=False, oneof=None):
    f = msg.field.add()
    f.name, f.number = name, number
    f.label = _T.LABEL_REPEATED if repeated else _T.LABEL_OPTIONAL
    if typ in _SCALAR:
        f.type = _SCALAR[typ]
    else:
        f.type, f.type_name = _T.TYPE_MESSAGE, typ
    if oneof is not None:
        f.oneof_index = oneof
    return f


def _map_field(msg, scope, name, number, value_type):
    """map<string, V> == repeated nested <Name>Entry{key=1, value=2} with map_entry set"""
    entry_name = "".join(p.capitalize() for p in name.split("_")) + "Entry"
    e = msg.nested_type.add()
    e.name = entry_name
    e.options.map_entry = True
    _field(e, "key", 1, "string")
    _field(e, "value", 2, value_type)
    _field(msg, name, number, f"{scope}.{entry_name}", repeated=True)


def _build():
    fd = descriptor_pb2.FileDescriptorProto()
    fd.name, fd.package, fd.syntax = "kserve_b200/grpc_predict_v2.proto", "inference", "proto3"

    def message(name):
        m = fd.message_type.add()
        m.name = name
        return m

    for name, fields in {
        "ServerLiveRequest": [], "ServerLiveResponse": [("live", 1, "bool")],
        "ServerReadyRequest": [], "ServerReadyResponse": [("ready", 1, "bool")],
        "ModelReadyRequest": [("name", 1, "string"), ("version", 2, "string")], "ModelReadyResponse": [("ready", 1, "bool")],
        "ServerMetadataRequest": [],
        "ModelMetadataRequest": [("name", 1, "string"), ("version", 2, "string")],
        "RepositoryModelLoadRequest": [("model_name", 1, "string")],
        "RepositoryModelLoadResponse": [("model_name", 1, "string"), ("isLoaded", 2, "bool")],
        "RepositoryModelUnloadRequest": [("model_name", 1, "string")],
        "RepositoryModelUnloadResponse": [("model_name", 1, "string"), ("isUnloaded", 2, "bool")],
    }.items():
        m = message(name)
        for f in fields:
            _field(m, *f)
    m = message("ServerMetadataResponse")
    _field(m, "name", 1, "string"); _field(m, "version", 2, "string"); _field(m, "extensions", 3, "string", repeated=True)
    m = message("ModelMetadataResponse")
    t = m.nested_type.add()
    t.name = "TensorMetadata"
    _field(t, "name", 1, "string"); _field(t, "datatype", 2, "string"); _field(t, "shape", 3, "int64", repeated=True)
    _field(m, "name", 1, "string"); _field(m, "versions", 2, "string", repeated=True); _field(m, "platform", 3, "string")
    _field(m, "inputs", 4, ".inference.ModelMetadataResponse.TensorMetadata", repeated=True)
    _field(m, "outputs", 5, ".inference.ModelMetadataResponse.TensorMetadata", repeated=True)

    m = message("InferParameter")
    m.oneof_decl.add().name = "parameter_choice"
    _field(m, "bool_param", 1, "bool", oneof=0); _field(m, "int64_param", 2, "int64", oneof=0)
    _field(m, "string_param", 3, "string", oneof=0)
    m = message("InferTensorContents")
    for i, (n, t_) in enumerate([("bool_contents", "bool"), ("int_contents", "int32"), ("int64_contents", "int64"),
                                 ("uint_contents", "uint32"), ("uint64_contents", "uint64"), ("fp32_contents", "float"),
                                 ("fp64_contents", "double"), ("bytes_contents", "bytes")], start=1):
        _field(m, n, i, t_, repeated=True)

    def tensor(parent, scope, name):
        t_ = parent.nested_type.add()
        t_.name = name
        _field(t_, "name", 1, "string"); _field(t_, "datatype", 2, "string"); _field(t_, "shape", 3, "int64", repeated=True)
        _map_field(t_, f"{scope}.{name}", "parameters", 4, ".inference.InferParameter")
        _field(t_, "contents", 5, ".inference.InferTensorContents")

    m = message("ModelInferRequest")
    tensor(m, ".inference.ModelInferRequest", "InferInputTensor")
    ro = m.nested_type.add()
    ro.name = "InferRequestedOutputTensor"
    _field(ro, "name", 1, "string")
    _map_field(ro, ".inference.ModelInferRequest.InferRequestedOutputTensor", "parameters", 2, ".inference.InferParameter")
    _field(m, "model_name", 1, "string"); _field(m, "model_version", 2, "string"); _field(m, "id", 3, "string")
    _map_field(m, ".inference.ModelInferRequest", "parameters", 4, ".inference.InferParameter")
    _field(m, "inputs", 5, ".inference.ModelInferRequest.InferInputTensor", repeated=True)
    _field(m, "outputs", 6, ".inference.ModelInferRequest.InferRequestedOutputTensor", repeated=True)
    _field(m, "raw_input_contents", 7, "bytes", repeated=True)

    m = message("ModelInferResponse")
    tensor(m, ".inference.ModelInferResponse", "InferOutputTensor")
    _field(m, "model_name", 1, "string"); _field(m, "model_version", 2, "string"); _field(m, "id", 3, "string")
    _map_field(m, ".inference.ModelInferResponse", "parameters", 4, ".inference.InferParameter")
    _field(m, "outputs", 5, ".inference.ModelInferResponse.InferOutputTensor", repeated=True)
    _field(m, "raw_output_contents", 6, "bytes", repeated=True)

    svc = fd.service.add()
    svc.name = "GRPCInferenceService"
    for rpc in ("ServerLive", "ServerReady", "ModelReady", "ServerMetadata", "ModelMetadata", "ModelInfer",
                "RepositoryModelLoad", "RepositoryModelUnload"):
        r = svc.method.add()
        r.name, r.input_type, r.output_type = rpc, f".inference.{rpc}Request", f".inference.{rpc}Response"
    pool = descriptor_pool.DescriptorPool()
    return pool.Add(fd) or pool.FindFileByName(fd.name)


DESCRIPTOR = _build()
SERVICE_NAME = "inference.GRPCInferenceService"
METHODS = [m.name for m in DESCRIPTOR.services_by_name["GRPCInferenceService"].methods]
for _name, _desc in DESCRIPTOR.message_types_by_name.items():
    globals()[_name] = message_factory.GetMessageClass(_desc)
