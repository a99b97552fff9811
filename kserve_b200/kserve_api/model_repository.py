"""name -> model registry (mirrors python/kserve/kserve/model_repository.py:29-89)."""
from typing import Dict, Optional

from .model import BaseKServeModel


class ModelRepository:
    def __init__(self, models_dir: str = "/mnt/models"):
        self.models: Dict[str, BaseKServeModel] = {}
        self.models_dir = models_dir

    def set_models_dir(self, models_dir):
        self.models_dir = models_dir

    def get_model(self, name: str) -> Optional[BaseKServeModel]:
        return self.models.get(name)

    def get_models(self) -> Dict[str, BaseKServeModel]:
        return self.models

    async def is_model_ready(self, name: str) -> bool:
        m = self.get_model(name)
        if m is None:
            return False
        return await m.healthy()

    def update(self, model: BaseKServeModel, name: Optional[str] = None):
        self.models[name or model.name] = model

    def load(self, name: str) -> bool:
        m = self.get_model(name)
        return bool(m and m.load())

    def unload(self, name: str):
        if name in self.models:
            self.models[name].stop()
            del self.models[name]
        else:
            raise KeyError(f"model {name} does not exist")
