"""Open Inference Protocol (V2) tensor envelope: InferInput / InferRequest / InferOutput / InferResponse.

Mirrors the public surface of python/kserve/kserve/protocol/infer_type.py (constructors, ``from_bytes``
binary extension :593-666, ``as_numpy`` :235-258, ``to_rest`` :717, :1328) for the REST legs the LLM path
uses.  One deliberate difference: raw binary tensors stay a zero-copy
``np.frombuffer`` view — the reference converts them to Python lists (``set_data_from_numpy(...,
binary_data=False)`` :642-644), a full round trip this runtime skips so ``input_ids`` can go straight to
pinned memory.
"""
from __future__ import annotations

import json
import struct
import uuid
from typing import Any, Dict, List, Optional, Tuple, Union

import numpy as np

from ..errors import InvalidInput

_DT = {
    "BOOL": np.bool_, "UINT8": np.uint8, "UINT16": np.uint16, "UINT32": np.uint32, "UINT64": np.uint64,
    "INT8": np.int8, "INT16": np.int16, "INT32": np.int32, "INT64": np.int64,
    "FP16": np.float16, "FP32": np.float32, "FP64": np.float64, "BYTES": np.object_,
}
_NP2DT = {np.dtype(v): k for k, v in _DT.items() if k != "BYTES"}


def to_np_dtype(datatype: str):
    if datatype not in _DT:
        raise InvalidInput(f"unsupported datatype {datatype}")
    return _DT[datatype]


def from_np_dtype(dt) -> str:
    dt = np.dtype(dt)
    if dt == np.object_ or dt.type in (np.bytes_, np.str_):
        return "BYTES"
    if dt not in _NP2DT:
        raise InvalidInput(f"unsupported numpy dtype {dt}")
    return _NP2DT[dt]


def _serialize_bytes(arr: np.ndarray) -> bytes:
    out = bytearray()
    for item in arr.reshape(-1):
        b = item if isinstance(item, (bytes, bytearray)) else str(item).encode("utf-8")
        out += struct.pack("<I", len(b)) + b
    return bytes(out)


def serialize_byte_tensor(arr: np.ndarray) -> np.ndarray:
    """BYTES elements are <uint32 little-endian length><bytes> back to back (OIP binary extension).  Returned the way the
    reference (and tritonclient) return it: a 0-d object array whose `.item()` is the byte string."""
    return np.asarray(_serialize_bytes(arr), dtype=np.object_)


def deserialize_bytes_tensor(raw: bytes) -> np.ndarray:
    items, off = [], 0
    while off < len(raw):
        (n,) = struct.unpack_from("<I", raw, off)
        off += 4
        items.append(bytes(raw[off:off + n]))
        off += n
    return np.array(items, dtype=np.object_)


class _Tensor:
    def __init__(self, name: str, shape: List[int], datatype: str, data=None, parameters: Optional[Dict] = None):
        self._name, self._shape, self._datatype = name, list(shape), datatype.upper()
        self._parameters = parameters if parameters is not None else {}
        self._data = data
        self._raw_data: Optional[bytes] = None

    name = property(lambda s: s._name)
    datatype = property(lambda s: s._datatype)

    @property
    def parameters(self):
        return self._parameters

    @parameters.setter
    def parameters(self, v):
        self._parameters = v if v is not None else {}

    @property
    def shape(self):
        return self._shape

    @shape.setter
    def shape(self, v):
        self._shape = list(v)

    @property
    def data(self):
        return self._data

    @data.setter
    def data(self, v):
        self._data = v

    def as_numpy(self) -> np.ndarray:
        dtype = to_np_dtype(self._datatype)
        if self._raw_data is not None:
            if self._datatype == "BYTES":
                return deserialize_bytes_tensor(self._raw_data).reshape(self._shape)
            return np.frombuffer(self._raw_data, dtype=dtype).reshape(self._shape)   # zero copy
        if self._data is None:
            raise InvalidInput(f"'data' field is missing for tensor '{self._name}'")
        if isinstance(self._data, np.ndarray):
            return self._data.reshape(self._shape)
        if self._datatype == "BYTES":
            return np.array([x.encode("utf-8") if isinstance(x, str) else x for x in _flatten(self._data)],
                            dtype=np.object_).reshape(self._shape)
        return np.asarray(self._data, dtype=dtype).reshape(self._shape)

    def __eq__(self, other):
        if not isinstance(other, _Tensor):
            return False
        return (self._name == other._name and self._shape == other._shape and self._datatype == other._datatype
                and self._parameters == other._parameters and _same(self._data, other._data)
                and _same(self._raw_data, other._raw_data))

    __hash__ = None

    def as_string(self) -> List[str]:
        return [x.decode("utf-8") if isinstance(x, (bytes, bytearray)) else str(x) for x in self.as_numpy().reshape(-1)]

    def set_data_from_numpy(self, arr: np.ndarray, binary_data: bool = True):
        if not isinstance(arr, np.ndarray):
            raise InvalidInput("input tensor must be a numpy array")
        dt = from_np_dtype(arr.dtype)
        if dt != self._datatype:
            raise InvalidInput(f"got unexpected datatype {dt} from numpy array, expected {self._datatype}")
        self._shape = list(arr.shape)
        if binary_data:
            self._data = None
            self._raw_data = _serialize_bytes(arr) if dt == "BYTES" else np.ascontiguousarray(arr).tobytes()
            self._parameters["binary_data_size"] = len(self._raw_data)
        else:
            self._raw_data = None
            self._parameters.pop("binary_data_size", None)
            if dt == "BYTES":
                self._data = [x.decode("utf-8") if isinstance(x, (bytes, bytearray)) else str(x) for x in arr.reshape(-1)]
            else:
                self._data = arr.reshape(-1).tolist()

    def _to_dict(self, binary: bool, raw_out: List[bytes]) -> Dict[str, Any]:
        d: Dict[str, Any] = {"name": self._name, "shape": self._shape, "datatype": self._datatype}
        params = dict(self._parameters)
        if params or binary:
            d["parameters"] = params          # key order of the reference's REST body: ..., parameters, data
        if binary:
            raw = self._raw_data
            if raw is None:
                arr = self.as_numpy()
                raw = _serialize_bytes(arr) if self._datatype == "BYTES" else np.ascontiguousarray(arr).tobytes()
            params["binary_data_size"] = len(raw)
            raw_out.append(raw)
        else:
            params.pop("binary_data_size", None)
            if self._datatype == "FP16":
                raise InvalidInput(f"Sending FP16 data via JSON is not supported. Please use the binary data format for {self._name}")
            if self._raw_data is not None or isinstance(self._data, np.ndarray):
                arr = self.as_numpy()
                d["data"] = ([x.decode("utf-8") if isinstance(x, (bytes, bytearray)) else x for x in arr.reshape(-1)]
                             if self._datatype == "BYTES" else arr.reshape(-1).tolist())
            else:
                if self._data is None:
                    raise InvalidInput(f"'data' field is missing for tensor '{self._name}'")
                d["data"] = self._data
        if not params:
            d.pop("parameters", None)
        return d


def _same(a, b) -> bool:
    if a is None or b is None:
        return a is None and b is None
    if isinstance(a, np.ndarray) or isinstance(b, np.ndarray):
        return np.array_equal(np.asarray(a), np.asarray(b))
    if isinstance(a, (bytes, bytearray, memoryview)) or isinstance(b, (bytes, bytearray, memoryview)):
        return bytes(a) == bytes(b)
    return a == b


def _flatten(x):
    if isinstance(x, (list, tuple)):
        for y in x:
            yield from _flatten(y)
    else:
        yield x


class InferInput(_Tensor):
    pass


class InferOutput(_Tensor):
    pass


class RequestedOutput:
    def __init__(self, name: str, parameters: Optional[Dict] = None):
        self.name, self.parameters = name, parameters or {}

    def __eq__(self, other):
        return isinstance(other, RequestedOutput) and self.name == other.name and self.parameters == other.parameters

    __hash__ = None

    @property
    def binary_data(self) -> Optional[bool]:
        return self.parameters.get("binary_data")


class InferRequest:
    def __init__(self, model_name: str, infer_inputs: List[InferInput], request_id: Optional[str] = None,
                 parameters: Optional[Dict] = None, request_outputs: Optional[List[RequestedOutput]] = None,
                 model_version: Optional[str] = None):
        self.id = request_id
        self.model_name = model_name
        self.model_version = model_version
        self.inputs = infer_inputs
        self.parameters = parameters or {}
        self.request_outputs = request_outputs

    def __eq__(self, other):
        if not isinstance(other, InferRequest):
            return False
        return (self.id == other.id and self.model_name == other.model_name and self.model_version == other.model_version
                and self.parameters == other.parameters and self.inputs == other.inputs
                and (self.request_outputs or None) == (other.request_outputs or None))

    __hash__ = None

    @classmethod
    def from_bytes(cls, req_bytes: bytes, json_length: int, model_name: str) -> "InferRequest":
        """infer_type.py:593-666 — JSON header followed by the raw little-endian tensors, in input order."""
        try:
            d = json.loads(req_bytes[:json_length])
        except json.JSONDecodeError as e:
            raise InvalidInput(f"Unrecognized request format: {e}")
        inputs, start = [], json_length
        view = memoryview(req_bytes)
        for inp in d.get("inputs", []):
            params = inp.get("parameters") or {}
            t = InferInput(inp["name"], inp["shape"], inp["datatype"], parameters=dict(params))
            if inp.get("data") is not None:
                if t.datatype == "FP16":
                    raise InvalidInput(f"Receiving FP16 data via JSON is not supported. Please use the binary data format for input {t.name}")
                t.data = inp["data"]
            elif "binary_data_size" in params:
                n = params["binary_data_size"]
                if n is None:
                    raise InvalidInput(f"'binary_data_size' is not specified for input '{t.name}' for model '{model_name}'")
                t._raw_data = view[start:start + n]
                start += n
            else:
                raise InvalidInput(f"'data' field is missing for input '{t.name}' for model '{model_name}'")
            inputs.append(t)
        outs = None
        if d.get("outputs") is not None:
            outs = [RequestedOutput(o["name"], o.get("parameters")) for o in d["outputs"]]
        return cls(model_name=model_name, request_id=d.get("id"), parameters=d.get("parameters"), infer_inputs=inputs,
                   request_outputs=outs)

    @classmethod
    def from_dict(cls, d: Dict, model_name: str) -> "InferRequest":
        raw = json.dumps(d, separators=(",", ":")).encode()
        return cls.from_bytes(raw, len(raw), model_name)

    def as_dataframe(self):
        """one column per input tensor, named after it (infer_type.py:849-864)"""
        import pandas as pd
        dfs = []
        for t in self.inputs:
            data = t.data if t._raw_data is None and t.data is not None else t.as_numpy().reshape(-1).tolist()
            if t.datatype == "BYTES":
                data = [str(v, "utf-8") if isinstance(v, (bytes, bytearray)) else v for v in data]
            dfs.append(pd.DataFrame(data, columns=[t.name]))
        return pd.concat(dfs, axis=1)

    def get_input_by_name(self, name: str) -> Optional[InferInput]:
        for i in self.inputs:
            if i.name == name:
                return i
        return None

    @property
    def use_binary_outputs(self) -> bool:
        if self.parameters.get("binary_data_output"):
            return True
        return any(o.binary_data for o in (self.request_outputs or []))

    def to_rest(self) -> Tuple[Union[bytes, Dict], Optional[int]]:
        raws: List[bytes] = []
        ins = [i._to_dict(i._raw_data is not None, raws) for i in self.inputs]
        res: Dict[str, Any] = {"id": self.id or str(uuid.uuid4()), "model_name": self.model_name, "inputs": ins}
        if self.parameters:
            res["parameters"] = self.parameters
        if self.request_outputs:
            res["outputs"] = [{"name": o.name, **({"parameters": o.parameters} if o.parameters else {})} for o in self.request_outputs]
        if raws:
            j = json.dumps(res, separators=(",", ":")).encode()
            return j + b"".join(bytes(r) for r in raws), len(j)
        return res, None


class InferResponse:
    def __init__(self, response_id: str, model_name: str, infer_outputs: List[InferOutput],
                 model_version: Optional[str] = None, parameters: Optional[Dict] = None,
                 use_binary_outputs: bool = False, requested_outputs: Optional[List[RequestedOutput]] = None):
        self.id = response_id
        self.model_name = model_name
        self.model_version = model_version
        self.outputs = infer_outputs
        self.parameters = parameters or {}
        self._use_binary_outputs = use_binary_outputs
        self._requested_outputs = requested_outputs

    def get_output_by_name(self, name: str) -> Optional[InferOutput]:
        for o in self.outputs:
            if o.name == name:
                return o
        return None

    def __eq__(self, other):
        if not isinstance(other, InferResponse):
            return False
        return (self.id == other.id and self.model_name == other.model_name and self.model_version == other.model_version
                and self.parameters == other.parameters and self.outputs == other.outputs)

    __hash__ = None

    @classmethod
    def from_bytes(cls, res_bytes: bytes, json_length: int) -> "InferResponse":
        """The client-side inverse of to_rest (infer_type.py `InferResponse.from_bytes`): JSON header + raw tensors in
        output order.  BYTES tensors come back as a list of str in `.data`, numeric ones as a numpy array."""
        try:
            d = json.loads(res_bytes[:json_length])
        except json.JSONDecodeError as e:
            raise InvalidInput(f"Unrecognized request format: {e}")
        outs, start = [], json_length
        for o in d.get("outputs", []):
            params = o.get("parameters") or {}
            t = InferOutput(o["name"], o["shape"], o["datatype"], parameters=dict(params))
            if o.get("data") is not None:
                t.data = o["data"]
            elif params.get("binary_data_size") is not None:
                n = params["binary_data_size"]
                raw = bytes(res_bytes[start:start + n])
                start += n
                if t.datatype == "BYTES":
                    t.data = [x.decode("utf-8") for x in deserialize_bytes_tensor(raw)]
                else:
                    t.data = np.frombuffer(raw, dtype=to_np_dtype(t.datatype)).reshape(t.shape)
            else:
                raise InvalidInput(f"'data' field is missing for output '{t.name}' for model '{d.get('model_name')}'")
            outs.append(t)
        return cls(response_id=d.get("id"), model_name=d.get("model_name"), infer_outputs=outs,
                   model_version=d.get("model_version"), parameters=d.get("parameters"))

    def to_rest(self) -> Tuple[Union[bytes, Dict], Optional[int]]:
        """infer_type.py:1328-1404: dict, or JSON header + raw tensors when binary outputs were requested."""
        want = {o.name: o for o in (self._requested_outputs or [])}
        raws: List[bytes] = []
        outs = []
        for o in self.outputs:
            if want and o.name not in want:
                continue
            binary = self._use_binary_outputs
            if o.name in want and want[o.name].binary_data is not None:
                binary = bool(want[o.name].binary_data)
            outs.append(o._to_dict(binary, raws))
        res: Dict[str, Any] = {"id": self.id, "model_name": self.model_name, "model_version": self.model_version, "outputs": outs}
        if self.parameters:
            res["parameters"] = self.parameters
        if raws:
            j = json.dumps(res, separators=(",", ":")).encode()
            return j + b"".join(bytes(r) for r in raws), len(j)
        return res, None


def _contains_fp16_datatype(infer_response: InferResponse) -> bool:
    """FP16 outputs force the binary response format (JSON cannot carry them)."""
    return any(o.datatype == "FP16" for o in infer_response.outputs)


def get_predict_input(payload: Union[Dict, InferRequest], columns: Optional[List] = None):
    """kserve/utils/utils.py:149-192 — V1 dict (`instances` / `inputs`) or V2 InferRequest -> what the model code consumes:
    a numpy array, a list of str (string instances / a BYTES tensor sent as JSON strings), or a pandas DataFrame (V1
    instances that are dicts; V2 requests with parameters.content_type == "pd")."""
    if isinstance(payload, dict):
        inst = payload["inputs"] if "inputs" in payload else payload["instances"]
        if len(inst) == 0:
            return np.array(inst)
        first = inst[0]
        if isinstance(first, dict) or (isinstance(first, list) and len(first) != 0 and isinstance(first[0], dict)):
            import pandas as pd
            return pd.concat([pd.DataFrame(i, columns=columns) for i in inst], axis=0)
        if isinstance(first, str):
            return inst
        return np.array(inst)
    if isinstance(payload, InferRequest):
        ct = (payload.parameters or {}).get("content_type")
        if ct is not None and not isinstance(ct, str):          # gRPC hands over an InferParameter
            ct = getattr(ct, "string_param", None)
        if ct == "pd":
            return payload.as_dataframe()
        t = payload.inputs[0]
        if t.datatype == "BYTES" and t._raw_data is None and isinstance(t.data, list) and len(t.data) > 0 and isinstance(t.data[0], str):
            return t.data
        return t.as_numpy()
    raise InvalidInput(f"unsupported payload type {type(payload)}")


def get_predict_response(payload, result, model_name: str):
    """kserve/utils/utils.py:194-254 — model result -> V1 dict or V2 InferResponse mirroring the request's encoding.  `result`:
    numpy array, list, list of str (-> one BYTES tensor), pandas DataFrame (one output per column), or — an extension the
    runtime model uses — a dict name -> array."""
    try:
        import pandas as pd
        is_df = isinstance(result, pd.DataFrame)
    except ImportError:                                          # pragma: no cover
        is_df = False
    if isinstance(payload, dict):
        if is_df:
            return {"predictions": [row.to_dict() for _, row in result.iterrows()]}
        return {"predictions": result.tolist() if isinstance(result, np.ndarray) else result}
    if not isinstance(payload, InferRequest):
        raise InvalidInput(f"unsupported payload type {type(payload)}")
    if is_df:
        items = [(col, result[col].to_numpy()) for col in result.columns]
    elif isinstance(result, dict):
        items = list(result.items())
    elif isinstance(result, list) and len(result) > 0 and isinstance(result[0], str):
        items = [("output-0", np.array(result, dtype=np.object_))]
    else:
        items = [("output-0", np.array(result) if isinstance(result, list) else result)]
    outs = []
    for name, arr in items:
        o = InferOutput(name, list(arr.shape), from_np_dtype(arr.dtype))
        o.data = arr
        outs.append(o)
    return InferResponse(response_id=payload.id or str(uuid.uuid4()), model_name=model_name, infer_outputs=outs,
                         use_binary_outputs=payload.use_binary_outputs, requested_outputs=payload.request_outputs)
