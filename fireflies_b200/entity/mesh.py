"""``fireflies/entity/mesh.py`` -- Mesh, same interface; vertex transforms run in libffb200."""
from __future__ import annotations

import os

import torch

from . import base
from .. import _native as nat
from .. import sampling
from ..utils import math as ffmath


class Mesh(base.Transformable):
    _KIND = nat.ENTITY_MESH

    def __init__(self, name: str, vertex_data: torch.Tensor, device: torch.device = torch.device("cuda")):
        super().__init__(name, device)
        self._vertices = vertex_data.to(self._device)
        self._vertices_animation = None
        ones = torch.ones(3, device=self._device)
        self._scale_sampler = sampling.UniformSampler(ones.clone(), ones.clone(), device=self._device)
        self._animated = False
        self._anim_data_train = None
        self._anim_data_eval = None
        self._animation_func = None
        self._animation_sampler = None

    def _samplers(self) -> list:
        return super()._samplers() + [self._scale_sampler]

    def set_scale_sampler(self, sampler) -> None:
        self._scale_sampler = sampler

    def scale_x(self, min_scale: float, max_scale: float) -> None:
        self._randomizable = True
        self.update_index_from_sampler(self._scale_sampler, min_scale, max_scale, 0)

    def scale_y(self, min_scale: float, max_scale: float) -> None:
        self._randomizable = True
        self.update_index_from_sampler(self._scale_sampler, min_scale, max_scale, 1)

    def scale_z(self, min_scale: float, max_scale: float) -> None:
        self._randomizable = True
        self.update_index_from_sampler(self._scale_sampler, min_scale, max_scale, 2)

    def scale(self, min: torch.Tensor, max: torch.Tensor) -> None:
        self._randomizable = True
        self._scale_sampler.set_sample_interval(min.to(self._device), max.to(self._device))

    def animated(self) -> bool:
        return self._animated

    def add_animation(self, animation_data: torch.Tensor) -> None:
        self._animation_vertices = animation_data.to(self._device)
        self._animated = True
        self._randomizable = True

    def add_animation_func(self, func, min_range, max_range) -> None:
        self._animation_func = func
        self._animation_sampler = sampling.UniformSampler(min_range, max_range, device=self._device)
        self._animated = True
        self._randomizable = True

    def add_train_animation(self, frames: torch.Tensor, min: int = None, max: int = None) -> None:
        """Tensor counterpart of ``add_train_animation_from_obj`` (``frames`` = ``[F,V,3]``), same bookkeeping
        (entity/mesh.py:75-91, including that ``min`` is ignored)."""
        self._anim_data_train = frames.to(self._device).float().contiguous()
        if self._animation_sampler:
            self._animation_sampler.set_train_interval(0, self._anim_data_train.shape[0] if max is None else max)
            return
        self._animation_sampler = sampling.AnimationSampler(0, 1, 0, 1, device=self._device)
        self._animation_sampler.set_train_interval(0, self._anim_data_train.shape[0] if max is None else max)
        self._animated = True

    def add_eval_animation(self, frames: torch.Tensor, min: int = None, max: int = None) -> None:
        """Tensor counterpart of ``add_eval_animation_from_obj`` (entity/mesh.py:93-109)."""
        self._anim_data_eval = frames.to(self._device).float().contiguous()
        if self._animation_sampler:
            self._animation_sampler.set_eval_interval(0, self._anim_data_eval.shape[0] if max is None else max)
            return
        self._animation_sampler = sampling.AnimationSampler(0, 1, 0, 1, device=self._device)
        self._animation_sampler.set_eval_interval(0, self._anim_data_eval.shape[0] if max is None else max)

    def add_train_animation_from_obj(self, path: str, min: int = None, max: int = None) -> None:
        self.add_train_animation(self.load_animation(path), min, max)

    def add_eval_animation_from_obj(self, path: str, min: int = None, max: int = None) -> None:
        self.add_eval_animation(self.load_animation(path), min, max)

    def train(self) -> None:
        super().train()
        if self._animation_sampler:
            self._animation_sampler.train()

    def eval(self) -> None:
        super().eval()
        if self._animation_sampler:
            self._animation_sampler.eval()

    def set_faces(self, faces: torch.Tensor) -> None:
        self._faces = faces.to(self._device)

    def set_vertices(self, vertices: torch.Tensor) -> None:
        self._vertices = vertices.to(self._device)

    def sample_scale(self) -> torch.Tensor:
        """entity/mesh.py:131-139."""
        random_scale = self._scale_sampler.sample()
        dev = self._world.device
        return self._compose_local(None, None, random_scale, torch.eye(4, device=dev), torch.zeros(3, device=dev),
                                   nat.ENTITY_MESH)

    def _draw_trs(self):
        t, r, _ = super()._draw_trs()
        self._sampled_scale = self._scale_sampler.sample()
        return t, r, self._sampled_scale

    def _sample_attributes(self) -> None:
        pass    # Mesh.randomize never samples float/vec3 attributes (entity/mesh.py:141-150)

    def faces(self) -> torch.Tensor:
        return self._faces

    def get_vertices(self) -> torch.Tensor:
        return self._vertices

    def get_randomized_vertices(self) -> torch.Tensor:
        """entity/mesh.py:158-165."""
        temp_vertex = self.sample_animation() if self._animated else self._vertices
        return ffmath.transform_points(temp_vertex, self.world())

    def load_animation(self, path: str) -> torch.Tensor:
        """entity/mesh.py:167-181 with a minimal OBJ vertex reader (pywavefront is not a dependency)."""
        animation_data = []
        for file in sorted(os.listdir(path)):
            if file.endswith(".obj"):
                verts = []
                with open(os.path.join(path, file)) as fh:
                    for line in fh:
                        if line.startswith("v "):
                            verts.append([float(x) for x in line.split()[1:4]])
                animation_data.append(torch.tensor(verts, device=self._device).reshape(-1, 3))
        return torch.stack(animation_data)

    def sample_animation(self):
        """entity/mesh.py:183-198."""
        if not self._animated:
            return self._vertices
        time_sample = self._animation_sampler.sample()
        if self._animation_func is not None:
            return self._animation_func(self._vertices, time_sample)
        elif self._anim_data_train is not None and self._anim_data_eval is not None:
            return self._anim_data_train[time_sample] if self._train else self._anim_data_eval[time_sample]
        return None
