"""``fireflies/scene.py`` -- Scene facade with the reference's discovery / train / eval / randomize API.

Mitsuba stays the external renderer: this module only reads and writes the ``mi.traverse`` parameter map.  When
``mitsuba`` is not importable (tests, benches) values are written back as plain torch tensors / python lists, or
through the constructors a fake parameter object exposes as ``mitsuba_params.types`` (tests/fake_mitsuba.py).
``Scene.batch()`` returns the B200-native batched randomiser (fireflies_b200/batch.py).
"""
from __future__ import annotations

from typing import List

import torch

from . import emitter, entity, material

try:                                     # pragma: no cover - mitsuba is not part of this image
    import mitsuba as mi
except Exception:                        # noqa: BLE001
    mi = None


class _Passthrough:
    Float32 = staticmethod(lambda x: x)
    Transform4f = staticmethod(lambda x: x)
    TensorXf = staticmethod(lambda x: x)


class Scene:
    MESH_KEYS = ["mesh", "ply"]
    CAM_KEYS = ["camera", "perspective", "perspectivecamera"]
    PROJ_KEYS = ["projector"]
    MAT_KEYS = ["mat", "bsdf"]
    LIGHT_KEYS = ["light", "spot"]
    TEX_KEYS = ["tex"]

    def __init__(self, mitsuba_params, device: torch.device = torch.device("cuda")):
        self._meshes = []
        self._projector = None
        self._camera = None
        self._lights = []
        self._curves = []
        self._materials = []
        self._transformables = []
        self._device = device
        self._mitsuba_params = mitsuba_params
        self._types = getattr(mitsuba_params, "types", None) or (mi if mi is not None else _Passthrough)
        self._train = True
        self.init_from_params(self._mitsuba_params)

    def device(self):
        return self._device

    # ---- accessors (scene.py:44-90) ------------------------------------------------------------------
    def mesh_at(self, index: int):
        return self._meshes[index]

    def meshes(self):
        return self._meshes

    def get_mesh(self, name: str):
        return next((m for m in self._meshes if m.name() == name), None)

    def mesh(self, name: str):
        return self.get_mesh(name)

    def light_at(self, index: int):
        return self._lights[index]

    def lights(self):
        return self._lights

    def get_light(self, name: str):
        return next((m for m in self._lights if m.name() == name), None)

    def light(self, name: str):
        return self.get_light(name)

    def material_at(self, index: int):
        return self._materials[index]

    def materials(self):
        return self._materials

    def get_material(self, name: str):
        return next((m for m in self._materials if m.name() == name), None)

    def material(self, name: str):
        return self.get_material(name)

    # ---- discovery (scene.py:92-201) -----------------------------------------------------------------
    def init_from_params(self, mitsuba_params) -> None:
        param_keys = sorted(set(key.split(".")[0] for key in mitsuba_params.keys()))
        for key in param_keys:
            low = key.lower()
            if any(k in low for k in self.MESH_KEYS):
                self.load_mesh(key)
            elif any(k in low for k in self.CAM_KEYS):
                self.load_camera(key)
            elif any(k in low for k in self.PROJ_KEYS):
                self.load_projector(key)
            elif any(k in low for k in self.LIGHT_KEYS):
                self.load_light(key)
            elif any(k in low for k in self.MAT_KEYS):
                self.load_material(key)

    @staticmethod
    def _as_tensor(value, device) -> torch.Tensor:
        if hasattr(value, "torch"):
            value = value.torch()
        return torch.as_tensor(value, dtype=torch.float32).to(device)

    def _to_world(self, base_key: str) -> torch.Tensor:
        m = self._mitsuba_params[base_key + ".to_world"].matrix
        m = m.torch() if hasattr(m, "torch") else torch.as_tensor(m)
        return m.squeeze().float().to(self._device)

    def load_mesh(self, base_key: str):
        vertices = self._as_tensor(self._mitsuba_params[base_key + ".vertex_positions"], self._device).reshape(-1, 3)
        centroid = vertices.sum(dim=0, keepdim=True) / vertices.shape[0]
        transformable_mesh = entity.Mesh(base_key, vertices - centroid, self._device)
        transformable_mesh.set_centroid(centroid)
        self._meshes.append(transformable_mesh)

    def load_camera(self, base_key: str) -> None:
        cam = entity.Transformable(base_key, self._device)
        cam.set_world(self._to_world(base_key))
        cam.set_randomizable(False)
        self._camera = cam

    def load_projector(self, base_key: str) -> None:
        proj = entity.Transformable(base_key, self._device)
        proj.set_world(self._to_world(base_key))
        proj.set_randomizable(False)
        self._projector = proj

    @staticmethod
    def _is_transform(value) -> bool:
        return hasattr(value, "matrix")

    @staticmethod
    def _is_scalar(value) -> bool:
        if isinstance(value, float):
            return True
        return mi is not None and isinstance(value, mi.Float) or bool(getattr(value, "is_scalar", False))

    def _register_attributes(self, ent, base_key: str) -> None:
        for key in [k for k in self._mitsuba_params.keys() if base_key in k]:
            key_without_base = ".".join(key.split(".")[1:])
            value = self._mitsuba_params[key]
            if self._is_transform(value):
                continue
            if self._is_scalar(value):
                v = float(value.value) if hasattr(value, "value") else float(value)
                ent.add_float_key(key_without_base, v, v)
            elif hasattr(value, "__len__") and len(value) == 3:
                v = self._as_tensor(value, self._device).squeeze()
                ent.add_vec3_key(key_without_base, v, v)

    def load_light(self, base_key: str) -> None:
        new_light = emitter.Light(base_key, device=self._device)
        if base_key + ".to_world" in self._mitsuba_params.keys():
            new_light.set_world(self._to_world(base_key))
        self._register_attributes(new_light, base_key)
        new_light.set_randomizable(False)
        self._lights.append(new_light)

    def load_material(self, base_key: str) -> None:
        new_material = material.Material(base_key, device=self._device)
        self._register_attributes(new_material, base_key)
        new_material.set_randomizable(False)
        self._materials.append(new_material)

    def _all(self):
        out = list(self._meshes) + list(self._lights) + list(self._materials)
        return out + [e for e in (self._camera, self._projector) if e is not None]

    def train(self) -> None:
        self._train = True
        for e in self._all():
            e.train()

    def eval(self) -> None:
        self._train = False
        for e in self._all():
            e.eval()

    def load_curve(self, path: str, name: str = "Curve") -> None:
        """scene.py:237-241: a NURBS camera path from a Blender .obj.  (The reference calls ``fireflies.utils.importBlenderNurbsObj``
        -- the function lives in ``fireflies.utils.io`` -- and appends to ``self.curves``, an attribute that does not exist; the
        evident intent is implemented: ``utils.io.importBlenderNurbsObj`` and ``self._curves``.)"""
        from .utils import io as ffio
        curve = ffio.importBlenderNurbsObj(path, device=self._device)
        self._curves.append(entity.Curve(name, curve, self._device))

    def curves(self) -> list:
        return self._curves

    # ---- write-back to Mitsuba (scene.py:243-342) ----------------------------------------------------
    def update_meshes(self) -> None:
        for mesh in self._meshes:
            if not mesh.randomizable():
                continue
            vertex_data = mesh.get_randomized_vertices()
            self._mitsuba_params[mesh.name() + ".vertex_positions"] = self._types.Float32(vertex_data.flatten())

    def _write_entity(self, ent, write_world: bool) -> None:
        name = ent.name()
        if write_world:
            self._mitsuba_params[name + ".to_world"] = self._types.Transform4f(ent.world().tolist())
        for key, value in ent.get_randomized_float_attributes().items():
            joined = name + "." + key
            temp_type = type(self._mitsuba_params[joined])
            self._mitsuba_params[joined] = temp_type(value.item())
        for key, value in ent.get_randomized_vec3_attributes().items():
            joined = name + "." + key
            temp_type = type(self._mitsuba_params[joined])
            self._mitsuba_params[joined] = temp_type(value.tolist())

    def update_camera(self) -> None:
        if self._camera.randomizable():
            self._write_entity(self._camera, True)

    def update_projector(self) -> None:
        if self._projector.randomizable():
            self._write_entity(self._projector, True)

    def update_lights(self) -> None:
        for light in self._lights:
            if light.randomizable():
                self._write_entity(light, light.name() + ".to_world" in self._mitsuba_params.keys())

    def update_materials(self) -> None:
        for mat in self._materials:
            if mat.randomizable():
                self._write_entity(mat, False)

    def randomize_list(self, entity_list: List) -> None:
        """scene.py:344-358: roots first, then down each single-child chain."""
        for ent in [e for e in entity_list if e.parent() is None]:
            ent.randomize()
            it = ent.child()
            while it is not None:
                it.randomize()
                it = it.child()

    def randomize(self) -> None:
        """scene.py:360-384."""
        self.randomize_list(self._meshes)
        self.randomize_list(self._lights)
        self.randomize_list(self._materials)
        if self._camera is not None:
            self._camera.randomize()
        if self._projector is not None:
            self._projector.randomize()
        self.update_meshes()
        if self._camera is not None:
            self.update_camera()
        if self._projector is not None:
            self.update_projector()
        self.update_lights()
        self.update_materials()
        self._mitsuba_params.update()

    # ---- B200-native batched path ------------------------------------------------------------------------
    def batch(self, seed: int = 0):
        """Batched randomiser over this scene's entities: B samples per launch, counter-based RNG."""
        from .batch import SceneBatch
        return SceneBatch(self, seed=seed)
