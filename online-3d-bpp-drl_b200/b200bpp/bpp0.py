"""Single-env view with the reference's gym surface (envs/bpp0/bin3D.py:9-127, space.py:19-181,
binCreator.py:5-72) for the callers that drive ONE bin from the host: unified_test.py, acktr/reorder.py (deep
copies the env and assigns box_creator.box_list, reorder.py:164), MCTS/node.py.

The item list lives on the host (it must be assignable and previewable); every feasibility check, placement and
mask is computed on the GPU through the same C ABI as the batched env (a 1-bin env used as the compute engine:
state in, one step, state out).  This is plumbing for BASELINE config 1, not a fast path -- rollouts belong to
BinPackingVecEnv.
"""
import copy

import numpy as np
import torch

from .spaces import Box as _BoxSpace, Discrete
from .vec_env import BinPackingVecEnv, default_item_set


class Box(object):
    """space.py:6-16"""

    def __init__(self, x, y, z, lx, ly, lz):
        self.x, self.y, self.z, self.lx, self.ly, self.lz = x, y, z, lx, ly, lz

    def standardize(self):
        return (self.x, self.y, self.z, self.lx, self.ly, self.lz)


class BoxCreator(object):
    """binCreator.py:5-22"""

    def __init__(self):
        self.box_list = []

    def reset(self):
        self.box_list.clear()

    def generate_box_size(self, **kwargs):
        pass

    def preview(self, length):
        while len(self.box_list) < length:
            self.generate_box_size()
        return copy.deepcopy(self.box_list[:length])

    def drop_box(self):
        self.box_list.pop(0)


class RandomBoxCreator(BoxCreator):
    """binCreator.py:24-40 (NumPy global RNG, like the reference)"""

    # binCreator.py:25-29: sides 2..6, 125 items (the trainer passes args.box_size_set, sides 2..5, instead)
    default_box_set = [(2 + i, 2 + j, 2 + k) for i in range(5) for j in range(5) for k in range(5)]

    def __init__(self, box_size_set=None):
        super().__init__()
        self.box_set = box_size_set if box_size_set is not None else RandomBoxCreator.default_box_set

    def generate_box_size(self, **kwargs):
        self.box_list.append(tuple(self.box_set[np.random.randint(0, len(self.box_set))]))


class LoadBoxCreator(BoxCreator):
    """binCreator.py:42-72: trajectory index += 1 per reset, (10,10,10) sentinel past the end."""

    def __init__(self, data_name=None, trajectories=None):
        super().__init__()
        self.data_name = data_name
        self.box_trajs = trajectories if trajectories is not None else torch.load(data_name)
        self.index = 0
        self.box_index = 0
        self.traj_nums = len(self.box_trajs)

    def reset(self, index=None):
        self.box_list.clear()
        self.recorder = []
        self.index = self.index + 1 if index is None else index
        self.boxes = [list(b) for b in self.box_trajs[self.index]]
        self.box_index = 0
        self.box_set = self.boxes
        self.box_set.append([10, 10, 10])

    def generate_box_size(self, **kwargs):
        item = self.box_set[self.box_index] if self.box_index < len(self.box_set) else (10, 10, 10)
        self.box_list.append(item)
        self.recorder.append(item)
        self.box_index += 1


class DeviceCutBoxCreator(BoxCreator):
    """--item-seq cut1 / cut2 for the single-env view: every reset() takes the next episode that the device generator
    (csrc/bpp_gen.cu; CuttingBoxCreator cutCreator.py:32-128 / MDlayerBoxCreator mdCreator.py:147-166) cuts for a 1-bin
    env; past its end the reference's sentinel follows (bin size for cut1, (10,10,10) for cut2)."""

    def __init__(self, kind, container_size, box_set, enable_rotation=False, seed=0, device=None):
        super().__init__()
        self.kind, self.bin_size = kind, tuple(int(v) for v in container_size)
        self._args = dict(container_size=self.bin_size, box_set=box_set, enable_rotation=enable_rotation, data_type=kind,
                          seed=seed, device=device, host_infos=False)
        self._env = None
        self.box_set, self.index = [], 0

    def __deepcopy__(self, memo):
        new = copy.copy(self)
        new._env = None                         # the copy replays its current episode; a reset would need its own stream
        new.box_list = copy.deepcopy(self.box_list, memo)
        new.box_set = list(self.box_set)
        return new

    def reset(self):
        self.box_list.clear()
        if self._env is None:
            self._env = BinPackingVecEnv(1, **self._args)
        self._env.reset()                       # an explicit reset moves the bin to its next episode
        self.box_set = self._env.episode_items(0)
        self.index = 0

    def generate_box_size(self, **kwargs):
        if self.index < len(self.box_set):
            self.box_list.append(self.box_set[self.index])
        else:
            self.box_list.append(self.bin_size if self.kind == "cut1" else (10, 10, 10))
        self.index += 1


class Space(object):
    """Host mirror of space.py:19-181; check_box / drop_box run on the device engine."""

    def __init__(self, game, width=10, length=10, height=10):
        self._game = game
        self.plain_size = np.array([width, length, height])
        self.plain = np.zeros((width, length), dtype=np.int32)
        self.boxes = []
        self.flags = []
        self.height = height

    def get_action_space(self):
        return self.plain_size[0] * self.plain_size[1]

    def get_ratio(self):
        vo = float(sum(b.x * b.y * b.z for b in self.boxes))
        return vo / float(self.plain_size[0] * self.plain_size[1] * self.plain_size[2])

    def idx_to_position(self, idx):
        return idx // self.plain_size[1], idx % self.plain_size[1]

    def position_to_index(self, position):
        assert len(position) == 2 and position[0] >= 0 and position[1] >= 0
        assert position[0] < self.plain_size[0] and position[1] < self.plain_size[1]
        return position[0] * self.plain_size[1] + position[1]

    def check_box(self, plain, x, y, lx, ly, z):
        from . import utils
        return utils.check_box(plain, x, y, lx, ly, z, tuple(int(v) for v in self.plain_size), variant=1)


class PackingGame(object):
    """envs/bpp0/bin3D.py:9-127 over one bin."""

    def __init__(self, box_creator=None, container_size=(20, 20, 20), box_set=None, data_name=None, test=False,
                 data_type="cut1", enable_rotation=False, device=None, **kwargs):
        self.box_creator = box_creator
        self.bin_size = tuple(int(v) for v in container_size)
        self.area = int(self.bin_size[0] * self.bin_size[1])
        self.can_rotate = bool(enable_rotation)
        self._device = device
        self._engine = None
        self.space = Space(self, *self.bin_size)
        if not test and box_creator is None:
            assert box_set is not None
            if data_type == "rs":
                self.box_creator = RandomBoxCreator(box_set)
            elif data_type in ("cut1", "cut2"):
                self.box_creator = DeviceCutBoxCreator(data_type, self.bin_size, box_set, self.can_rotate, device=device,
                                                       seed=kwargs.get("seed", 0))
            else:
                raise ValueError("data_type must be 'rs', 'cut1' or 'cut2'")
        if test:
            self.box_creator = LoadBoxCreator(data_name)
        self.act_len = self.area * (1 + self.can_rotate)
        self.obs_len = self.area * 4
        self.action_space = Discrete(self.act_len)
        self.observation_space = _BoxSpace(0.0, self.space.height, (self.obs_len,))

    # -- device engine ----------------------------------------------------------------------------------
    def _eng(self):
        if self._engine is None:
            dummy = (np.array([[[1, 1, 1]]], dtype=np.uint8), np.array([1], dtype=np.int32))
            self._engine = BinPackingVecEnv(1, self.bin_size, enable_rotation=self.can_rotate, trajectories=dummy,
                                            first_traj=0, mask_variant=1, device=self._device, host_infos=False)
            self._engine.reset()
        return self._engine

    def _push(self, plain=None):
        e = self._eng()
        e.import_state(plain=np.asarray(self.space.plain if plain is None else plain).reshape(1, -1),
                       box=np.asarray(self.next_box, dtype=np.int32).reshape(1, 3),
                       counter=[len(self.space.boxes)], vol=[int(sum(b.x * b.y * b.z for b in self.space.boxes))])
        return e

    def __deepcopy__(self, memo):
        new = copy.copy(self)
        new._engine = None                      # device handle is per object; state is host-side
        new.box_creator = copy.deepcopy(self.box_creator, memo)
        new.space = copy.copy(self.space)
        new.space._game = new
        new.space.plain = self.space.plain.copy()
        new.space.boxes = list(self.space.boxes)
        new.space.flags = list(self.space.flags)
        return new

    # -- gym surface ------------------------------------------------------------------------------------
    def get_box_ratio(self):
        b = self.next_box
        return (b[0] * b[1] * b[2]) / (self.space.plain_size[0] * self.space.plain_size[1] * self.space.plain_size[2])

    def reset(self):
        self.box_creator.reset()
        self.space = Space(self, *self.bin_size)
        self.box_creator.generate_box_size()
        return self.cur_observation

    @property
    def cur_observation(self):
        b = self.next_box
        planes = [self.space.plain] + [np.full(self.space.plain.shape, v, dtype=np.int32) for v in b]
        return np.stack(planes).reshape(-1)

    @property
    def next_box(self):
        return self.box_creator.preview(1)[0]

    def get_possible_position(self, plain=None):
        """Env-side mask with Space.check_box (bin3D.py:72-93): int32 [L, W]."""
        e = self._push(plain)
        _, masks = e.observe()
        m = masks[0, :self.area].to(torch.int32).cpu().numpy()
        if self.can_rotate:                     # the reference's env-side mask never looks at the rotated half
            from . import utils
            m = utils.mask_from_plain(np.asarray(self.space.plain if plain is None else plain).reshape(1, -1),
                                      [self.next_box], self.bin_size, False, 1)[0].to(torch.int32).cpu().numpy()
        return m.reshape(self.space.plain.shape)

    def step(self, action):
        idx = action[0] if isinstance(action, (np.ndarray, list)) else action
        idx = int(idx)
        box = self.next_box
        e = self._push()
        _, reward, done, _ = e.step_device(torch.tensor([idx], dtype=torch.int64, device=e.device))
        if int(done[0]):                        # bin3D.py:108-112: state unchanged, episode over
            info = {"counter": len(self.space.boxes), "ratio": self.space.get_ratio(), "mask": np.ones(self.act_len)}
            return self.cur_observation, 0.0, True, info
        flag = idx > self.area                  # bin3D.py:102
        pos = idx - self.area if flag else idx
        lx, ly = int(pos // self.bin_size[1]), int(pos % self.bin_size[1])
        x, y = (box[1], box[0]) if flag else (box[0], box[1])
        lz = int(self.space.plain[lx:lx + x, ly:ly + y].max())
        self.space.boxes.append(Box(x, y, box[2], lx, ly, lz))
        self.space.flags.append(flag)
        plain, _, _, _ = e.export_state()       # the device's heightmap after update_height_graph
        self.space.plain = plain[0].cpu().numpy().reshape(self.space.plain.shape).astype(np.int32)
        box_ratio = (box[0] * box[1] * box[2]) / float(np.prod(self.bin_size))
        self.box_creator.drop_box()
        self.box_creator.generate_box_size()
        info = {"counter": len(self.space.boxes), "ratio": self.space.get_ratio()}
        return self.cur_observation, box_ratio * 10, False, info

    def close(self):
        if self._engine is not None:
            self._engine.close()
            self._engine = None


# -- gym-style registry: main.py:223-227 registers 'Bpp-v0' -> envs.bpp0:PackingGame and calls gym.make(id, **kwargs) ------
_REGISTRY = {"Bpp-v0": PackingGame}


def register(id, entry_point=PackingGame, **kwargs):
    _REGISTRY[id] = entry_point


def make(id, **kwargs):
    return _REGISTRY[id](**kwargs)
