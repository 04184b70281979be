"""``fireflies/entity/base.py`` -- Transformable, same interface; 4x4 composition runs in
``ffb_compose_world`` (csrc/ffb_scene.cu) instead of ~190 tiny aten ops + 12 host syncs per entity."""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Sequence

import numpy as np
import torch

from .. import _native as nat
from .. import sampling

_ENT_WORDS = C.sizeof(nat.Entity) // 4        # 26
_ENT_INTS = 6


def compose_world(kinds: Sequence[int], parents: Sequence[int], randomizable: Sequence[int],
                  centroids: torch.Tensor, worlds: torch.Tensor, sampled: Optional[torch.Tensor]) -> torch.Tensor:
    """Low-level bridge to ``ffb_compose_world``.

    ``centroids`` f32 [E,3], ``worlds`` f32 [E,4,4] (device), ``sampled`` f32 [B,E,3,3] (per entity rows
    translation / rotation / scale) or None (B = 1, nothing randomizable).  Returns ``[B,E,4,4]``."""
    E = len(kinds)
    dev = worlds.device
    B = 1 if sampled is None else sampled.shape[0]
    ints = np.full((E, _ENT_INTS), -1, dtype=np.int32)
    for e in range(E):
        ints[e, 0], ints[e, 1], ints[e, 2] = kinds[e], parents[e], randomizable[e]
        if sampled is not None:
            ints[e, 3], ints[e, 4], ints[e, 5] = 3 * e, 3 * e + 1, 3 * e + 2
    table = torch.empty((E, _ENT_WORDS), dtype=torch.int32, device=dev)
    table[:, :_ENT_INTS] = torch.from_numpy(ints).to(dev)
    fv = table.view(torch.float32)
    fv[:, 6:9] = centroids.reshape(E, 3)
    fv[:, 9] = 0
    fv[:, 10:26] = worlds.reshape(E, 16)
    out = torch.empty((B, E, 4, 4), dtype=torch.float32, device=dev)
    S = 0 if sampled is None else 3 * E
    smp = None if sampled is None else nat.require_cuda(sampled.reshape(B, S, 3).contiguous(), torch.float32, "sampled")
    nat.check(nat.lib().ffb_compose_world(table.data_ptr(), E, B, nat.ptr(smp), S, out.data_ptr(), nat.stream()),
              "ffb_compose_world")
    nat.count()
    return out


class Transformable:
    _KIND = nat.ENTITY_PLAIN

    def __init__(self, name: str, device: torch.device = torch.device("cuda")):
        self._device = device
        self._name = name
        self._randomizable = False
        self._parent = None
        self._child = None
        self._train = True
        self._float_attributes = {}
        self._randomized_float_attributes = {}
        self._vec3_attributes = {}
        self._randomized_vec3_attributes = {}

        zeros = torch.zeros(3, device=self._device)
        self._rotation_sampler = sampling.UniformSampler(zeros.clone(), zeros.clone(), device=self._device)
        self._translation_sampler = sampling.UniformSampler(zeros.clone(), zeros.clone(), device=self._device)

        self._world = torch.eye(4, device=self._device)
        self._randomized_world = torch.eye(4, device=self._device)
        self._centroid_mat = torch.zeros((4, 4), device=self._device)
        self._eval_delta = 0.01
        self._num_updates = 0

    # ---- flags / attributes (entity/base.py:48-100) ----------------------------------------------
    def randomizable(self) -> bool:
        return self._randomizable

    def set_centroid(self, centroid: torch.Tensor) -> None:
        c = centroid.to(self._centroid_mat.device).reshape(-1)
        self._centroid_mat[0:3, 3] = c[0:3]

    def set_randomizable(self, randomizable: bool) -> None:
        self._randomizable = randomizable

    def get_randomized_vec3_attributes(self) -> dict:
        return self._randomized_vec3_attributes

    def get_randomized_float_attributes(self) -> dict:
        return self._randomized_float_attributes

    def vec3_attributes(self) -> dict:
        return self._vec3_attributes

    def float_attributes(self) -> dict:
        return self._float_attributes

    def add_float_sampler(self, key: str, sampler) -> None:
        self._randomizable = True
        self._float_attributes[key] = sampler

    def add_float_key(self, key: str, min: float, max: float) -> None:
        self._randomizable = True
        self._float_attributes[key] = sampling.UniformSampler(min, max, device=self._device)

    def add_vec3_key(self, key: str, min: torch.Tensor, max: torch.Tensor) -> None:
        self._randomizable = True
        self._vec3_attributes[key] = sampling.UniformSampler(min, max, device=self._device)

    def add_vec3_sampler(self, key: str, sampler) -> None:
        self._randomizable = True
        self._vec3_attributes[key] = sampler

    def parent(self):
        return self._parent

    def child(self):
        return self._child

    def name(self):
        return self._name

    def _samplers(self) -> list:
        return [self._translation_sampler, self._rotation_sampler]

    def train(self) -> None:
        self._train = True
        for s in self._samplers() + list(self._float_attributes.values()) + list(self._vec3_attributes.values()):
            s.train()

    def eval(self) -> None:
        self._train = False
        for s in self._samplers() + list(self._float_attributes.values()) + list(self._vec3_attributes.values()):
            s.eval()

    def set_world(self, _origin: torch.Tensor) -> None:
        self._world = _origin
        self._randomized_world = self._world.clone()

    def setParent(self, parent) -> None:
        self._parent = parent
        parent.setChild(self)

    def setChild(self, child) -> None:
        self._child = child

    def set_rotation_sampler(self, sampler) -> None:
        self._rotation_sampler = sampler

    def set_translation_sampler(self, sampler) -> None:
        self._translation_sampler = sampler

    def update_index_from_sampler(self, sampler, min, max, index) -> None:
        sampler.get_min()[index] = min          # in-place edit of the live range tensor (entity/base.py:141-146)
        sampler.get_max()[index] = max

    def rotate_x(self, min_rot: float, max_rot: float) -> None:
        self._randomizable = True
        self.update_index_from_sampler(self._rotation_sampler, min_rot, max_rot, 0)

    def rotate_y(self, min_rot: float, max_rot: float) -> None:
        self._randomizable = True
        self.update_index_from_sampler(self._rotation_sampler, min_rot, max_rot, 1)

    def rotate_z(self, min_rot: float, max_rot: float) -> None:
        self._randomizable = True
        self.update_index_from_sampler(self._rotation_sampler, min_rot, max_rot, 2)

    def rotate(self, min: torch.Tensor, max: torch.Tensor) -> None:
        self._randomizable = True
        self._rotation_sampler.set_sample_interval(min.to(self._device), max.to(self._device))

    def translate_x(self, min_translation: float, max_translation: float) -> None:
        self._randomizable = True
        self.update_index_from_sampler(self._translation_sampler, min_translation, max_translation, 0)

    def translate_y(self, min_translation: float, max_translation: float) -> None:
        self._randomizable = True
        self.update_index_from_sampler(self._translation_sampler, min_translation, max_translation, 1)

    def translate_z(self, min_translation: float, max_translation: float) -> None:
        self._randomizable = True
        self.update_index_from_sampler(self._translation_sampler, min_translation, max_translation, 2)

    def translate(self, min: torch.Tensor, max: torch.Tensor) -> None:
        self._randomizable = True
        self._translation_sampler.set_sample_interval(min.to(self._device), max.to(self._device))

    # ---- sampling + compose -------------------------------------------------------------------------
    def _compose_local(self, t, r, s, world, centroid, kind) -> torch.Tensor:
        dev = self._world.device
        zero, one = torch.zeros(3, device=dev), torch.ones(3, device=dev)
        smp = torch.stack([zero if t is None else t.reshape(3), zero if r is None else r.reshape(3),
                           one if s is None else s.reshape(3)]).reshape(1, 1, 3, 3)
        return compose_world([kind], [-1], [1], centroid.reshape(1, 3), world.reshape(1, 4, 4), smp)[0, 0]

    def _compose_randomized(self, t, r, s) -> torch.Tensor:
        """``randomize()``'s compose with this entity's own world / centroid: the one-row entity table stays on the device
        and is rebuilt only when the world matrix or the centroid changed (tensor identity + version), so a call is the
        sample stack and ONE launch instead of a table upload and a dozen slice writes."""
        dev = self._world.device
        # the key holds the tensors themselves (compared by identity): an id() alone can be recycled by a later tensor
        key = (self._world, self._world._version, self._centroid_mat, self._centroid_mat._version, self._KIND)
        cache = getattr(self, "_ent_cache", None)
        if cache is None or not (cache[0][0] is key[0] and cache[0][2] is key[2] and cache[0][1] == key[1] and cache[0][3] == key[3]
                                 and cache[0][4] == key[4]):
            ints = np.full((1, _ENT_INTS), -1, dtype=np.int32)
            ints[0, 0], ints[0, 1], ints[0, 2] = self._KIND, -1, 1
            ints[0, 3], ints[0, 4], ints[0, 5] = 0, 1, 2
            table = torch.zeros((1, _ENT_WORDS), dtype=torch.int32, device=dev)
            table[:, :_ENT_INTS] = torch.from_numpy(ints).to(dev)
            fv = table.view(torch.float32)
            fv[:, 6:9] = self._centroid_mat[0:3, 3].to(dev).float().reshape(1, 3)
            fv[:, 10:26] = self._world.to(dev).float().reshape(1, 16)
            consts = (torch.zeros(3, device=dev), torch.ones(3, device=dev))
            cache = (key, table, consts)
            self._ent_cache = cache
        _, table, (zero, one) = cache
        smp = torch.stack([zero if t is None else t.reshape(3), zero if r is None else r.reshape(3), one if s is None else s.reshape(3)])
        smp = nat.require_cuda(smp.float(), torch.float32, "sampled")
        out = torch.empty((1, 1, 4, 4), dtype=torch.float32, device=dev)
        nat.check(nat.lib().ffb_compose_world(table.data_ptr(), 1, 1, smp.data_ptr(), 3, out.data_ptr(), nat.stream()), "ffb_compose_world")
        nat.count()
        return out[0, 0]

    def sample_rotation(self) -> torch.Tensor:
        """entity/base.py:194-207: 4x4 of Pitch(r[2]) @ Yaw(r[1]) @ Roll(r[0])."""
        self._sampled_rotation = self._rotation_sampler.sample()
        dev = self._world.device
        return self._compose_local(None, self._sampled_rotation, None, torch.eye(4, device=dev), torch.zeros(3, device=dev),
                                   nat.ENTITY_PLAIN)

    def sample_translation(self) -> torch.Tensor:
        """entity/base.py:209-218."""
        self._random_translation = self._translation_sampler.sample()
        dev = self._world.device
        translation = self._compose_local(self._random_translation, None, None, torch.eye(4, device=dev),
                                          torch.zeros(3, device=dev), nat.ENTITY_PLAIN)
        self._last_translation = translation
        return translation

    def _draw_trs(self):
        """Draw order of the reference: translation, rotation (, scale) -- SURVEY.md App. B KAT5."""
        self._random_translation = self._translation_sampler.sample()
        self._sampled_rotation = self._rotation_sampler.sample()
        return self._random_translation, self._sampled_rotation, None

    def randomize(self) -> None:
        """entity/base.py:220-234: ``(T + C) @ R @ W`` in one compose launch, then attribute draws."""
        if not self.randomizable():
            return
        t, r, s = self._draw_trs()
        self._randomized_world = self._compose_randomized(t, r, s)
        self._sample_attributes()

    def _sample_attributes(self) -> None:
        for key, sampler in self._float_attributes.items():
            self._randomized_float_attributes[key] = sampler.sample()
        for key, sampler in self._vec3_attributes.items():
            self._randomized_vec3_attributes[key] = sampler.sample()

    def relative(self) -> bool:
        return self._parent is not None

    def _chain(self, attr: str) -> torch.Tensor:
        chain: List[Transformable] = []
        node = self
        while node is not None:
            chain.insert(0, node)
            node = node._parent
        if len(chain) == 1:
            return getattr(self, attr).clone()
        worlds = torch.stack([getattr(n, attr).to(self._world.device).float() for n in chain])
        E = len(chain)
        out = compose_world([nat.ENTITY_PLAIN] * E, list(range(-1, E - 1)), [0] * E,
                            torch.zeros(E, 3, device=worlds.device), worlds, None)
        return out[0, E - 1]

    def world(self) -> torch.Tensor:
        """entity/base.py:239-244: ``parent.world() @ randomized_world`` up the chain (one launch)."""
        return self._chain("_randomized_world")

    def nonRandomizedWorld(self) -> torch.Tensor:
        if self._parent is None:
            return self._world
        return self._chain("_world")
