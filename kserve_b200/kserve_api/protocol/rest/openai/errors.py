"""OpenAI-shaped errors (mirrors python/kserve/kserve/protocol/rest/openai/errors.py:22-66):
raise OpenAIError(str | ErrorResponse) -> JSON {"error": {code, message, param, type}} with the embedded
status code (500 for a bare string)."""
from http import HTTPStatus
from typing import Union

from fastapi.responses import JSONResponse

from .types import Error, ErrorResponse


class OpenAIError(Exception):
    def __init__(self, response: Union[str, ErrorResponse]):
        self.response = response

    def __str__(self):
        return self.response.error.message if isinstance(self.response, ErrorResponse) else self.response


def create_error_response(message: str, err_type: str = "BadRequestError", param: str = "",
                          status_code: HTTPStatus = HTTPStatus.BAD_REQUEST) -> ErrorResponse:
    return ErrorResponse(error=Error(message=message, type=err_type, param=param, code=str(status_code.value)))


async def openai_error_handler(_, exc: OpenAIError):
    response = exc.response if isinstance(exc.response, ErrorResponse) else create_error_response(
        message=str(exc), err_type=type(exc).__name__, status_code=HTTPStatus.INTERNAL_SERVER_ERROR)
    return JSONResponse(status_code=int(response.error.code), content=response.model_dump())
