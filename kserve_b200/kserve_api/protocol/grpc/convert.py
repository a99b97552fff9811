"""ModelInferRequest -> InferRequest and InferResponse -> ModelInferResponse
(what `InferRequest.from_grpc` / `InferResponse.to_grpc` do in python/kserve/kserve/protocol/infer_type.py:548-590, :1400-1465)."""
from typing import Any, Dict, List

import numpy as np

from ...errors import InvalidInput
from ..infer_type import (InferInput, InferRequest, InferResponse, RequestedOutput, _contains_fp16_datatype, _serialize_bytes,
                          to_np_dtype)
from . import pb

# datatype -> field of InferTensorContents (the OIP table)
CONTENT_FIELD = {"BOOL": "bool_contents", "INT8": "int_contents", "INT16": "int_contents", "INT32": "int_contents",
                 "INT64": "int64_contents", "UINT8": "uint_contents", "UINT16": "uint_contents", "UINT32": "uint_contents",
                 "UINT64": "uint64_contents", "FP32": "fp32_contents", "FP64": "fp64_contents", "BYTES": "bytes_contents"}


def to_http_parameters(params) -> Dict[str, Any]:
    out = {}
    for k, v in params.items():
        which = v.WhichOneof("parameter_choice")
        if which is not None:
            out[k] = getattr(v, which)
    return out


def to_grpc_parameters(params: Dict[str, Any]) -> Dict[str, Any]:
    out = {}
    for k, v in (params or {}).items():
        if isinstance(v, bool):
            out[k] = pb.InferParameter(bool_param=v)
        elif isinstance(v, int):
            out[k] = pb.InferParameter(int64_param=v)
        elif isinstance(v, str):
            out[k] = pb.InferParameter(string_param=v)
        else:
            raise InvalidInput(f"to_grpc: invalid parameter value: {v}")
    return out


def infer_request_from_grpc(req) -> InferRequest:
    raw = list(req.raw_input_contents)
    if raw and len(raw) != len(req.inputs):            # servicer.py:37-50
        raise InvalidInput(f"the number of inputs ({len(req.inputs)}) does not match the expected number of "
                           f"raw input contents ({len(raw)}) for model '{req.model_name}'.")
    inputs: List[InferInput] = []
    for i, t in enumerate(req.inputs):
        dt = t.datatype.upper()
        tin = InferInput(t.name, list(t.shape), dt, parameters=to_http_parameters(t.parameters))
        if raw:
            if t.HasField("contents"):
                raise InvalidInput(f"contents field must not be specified when using raw_input_contents for input "
                                   f"'{t.name}' for model '{req.model_name}'")
            tin._raw_data = raw[i]
        else:
            field = CONTENT_FIELD.get(dt)
            if field is None:
                raise InvalidInput(f"'{dt}' tensors must be sent as raw_input_contents (input '{t.name}')")
            tin.data = list(getattr(t.contents, field))
        inputs.append(tin)
    outs = [RequestedOutput(o.name, to_http_parameters(o.parameters)) for o in req.outputs] or None
    r = InferRequest(model_name=req.model_name, infer_inputs=inputs, request_id=req.id or None,
                     parameters=to_http_parameters(req.parameters), request_outputs=outs, model_version=req.model_version or None)
    r.from_grpc = True
    r.use_raw = bool(raw)
    return r


def infer_response_to_grpc(res: InferResponse):
    use_raw = res._use_binary_outputs or _contains_fp16_datatype(res)
    outs, raws = [], []
    for o in res.outputs:
        t: Dict[str, Any] = {"name": o.name, "datatype": o.datatype, "shape": o.shape}
        if o.parameters:
            t["parameters"] = to_grpc_parameters({k: v for k, v in o.parameters.items() if k != "binary_data_size"})
        raw = o._raw_data
        if raw is None and (use_raw or isinstance(o.data, np.ndarray)) and o.data is not None:
            arr = o.as_numpy()
            if use_raw:
                raw = _serialize_bytes(arr) if o.datatype == "BYTES" else np.ascontiguousarray(arr.astype(to_np_dtype(o.datatype), copy=False)).tobytes()
        if raw is not None and use_raw:
            raws.append(bytes(raw))
        else:
            field = CONTENT_FIELD.get(o.datatype)
            if field is None:
                raise InvalidInput("to_grpc: invalid output datatype")
            vals = o.as_numpy().reshape(-1).tolist() if (isinstance(o.data, np.ndarray) or o._raw_data is not None) else list(o.data)
            if o.datatype == "BYTES":
                vals = [v.encode("utf-8") if isinstance(v, str) else bytes(v) for v in vals]
            t["contents"] = {field: vals}
        outs.append(t)
    return pb.ModelInferResponse(id=res.id or "", model_name=res.model_name or "", model_version=res.model_version or "",
                                 outputs=outs, raw_output_contents=raws,
                                 parameters=to_grpc_parameters({k: v for k, v in (res.parameters or {}).items()
                                                                if isinstance(v, (bool, int, str))}))
