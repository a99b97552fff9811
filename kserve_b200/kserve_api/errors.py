"""kserve-native errors and their HTTP mapping (mirrors python/kserve/kserve/errors.py:23-186:
InvalidInput -> 400, ModelNotFound -> 404, ModelNotReady -> 503, InferenceError -> 500, NotImplemented -> 501,
bodies are {"error": str})."""
from http import HTTPStatus

from fastapi.responses import JSONResponse


class InferenceError(RuntimeError):
    def __init__(self, reason, status=None, debug_details=None):
        self.reason, self.status, self.debug_details = reason, status, debug_details

    def __str__(self):
        return self.reason


class InvalidInput(ValueError):
    def __init__(self, reason):
        self.reason = reason

    def __str__(self):
        return self.reason


class ModelNotFound(Exception):
    def __init__(self, model_name=None):
        self.reason = f"Model with name {model_name} does not exist."

    def __str__(self):
        return self.reason


class ModelNotReady(RuntimeError):
    def __init__(self, model_name: str, detail: str = None):
        self.model_name = model_name
        self.error_msg = f"Model with name {self.model_name} is not ready."
        if detail:
            self.error_msg = self.error_msg + " " + detail

    def __str__(self):
        return self.error_msg


class ServerNotReady(RuntimeError):
    def __str__(self):
        return "Server is not ready."


class ServerNotLive(RuntimeError):
    def __str__(self):
        return "Server is not live."


class UnsupportedProtocol(Exception):
    def __init__(self, protocol_version=None):
        self.reason = f"Unsupported protocol {protocol_version}."

    def __str__(self):
        return self.reason


async def invalid_input_handler(_, exc):
    return JSONResponse(status_code=HTTPStatus.BAD_REQUEST, content={"error": str(exc)})


async def inference_error_handler(_, exc):
    return JSONResponse(status_code=HTTPStatus.INTERNAL_SERVER_ERROR, content={"error": str(exc)})


async def generic_exception_handler(_, exc):
    return JSONResponse(status_code=HTTPStatus.INTERNAL_SERVER_ERROR, content={"error": f"{type(exc).__name__} : {str(exc)}"})


async def model_not_found_handler(_, exc):
    return JSONResponse(status_code=HTTPStatus.NOT_FOUND, content={"error": str(exc)})


async def model_not_ready_handler(_, exc):
    return JSONResponse(status_code=HTTPStatus.SERVICE_UNAVAILABLE, content={"error": str(exc)})


async def not_implemented_error_handler(_, exc):
    return JSONResponse(status_code=HTTPStatus.NOT_IMPLEMENTED, content={"error": str(exc)})


async def unsupported_protocol_error_handler(_, exc):
    return JSONResponse(status_code=HTTPStatus.NOT_IMPLEMENTED, content={"error": str(exc)})


async def server_not_ready_handler(_, exc):
    return JSONResponse(status_code=HTTPStatus.SERVICE_UNAVAILABLE, content={"error": str(exc)})
