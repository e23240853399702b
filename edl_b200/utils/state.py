"""Elastic train state: what must survive a membership change
(reference: python/edl/utils/state.py:25-217 -- partly unfinished there; the *intent* is built here).

* ``DataCheckpoint``  -- which record ranges of which files were already consumed (data-position resume)
* ``EpochAttr`` / ``TrainStatus`` -- epoch / global step counters and per-epoch world size, step time
* ``State``           -- the above + total batch size + user-defined serialisable + adjust callbacks
                         (``register_adjust_function``) that are invoked with the old and new world
                         size on every stage change (LR / batch rescale policies, see
                         ``linear_scale_lr`` / ``keep_total_batch``)
* ``TorchState``      -- State bound to a model/optimizer/trainer (the reference's ``PaddleState``)

Persistence: JSON in the store at ``state/<name>``, written only by the leader through a
compare-and-put on ``rank/0`` (same guard as the reference, state.py:186-200).
"""

from . import constants, unique_name
from . import train_status as edl_train_status
from .error_utils import handle_errors_until_timeout
from .exceptions import EdlEtcdIOError, EdlTableError
from .json_serializable import Serializable
from .string_utils import bytes_to_string


class DataCheckpoint(Serializable):
    def __init__(self, reader_name=None, file_list=None, processed_data=None):
        self.reader_name = reader_name
        self.file_list = file_list
        # file_idx (str) -> [[record_begin, record_end], ...]  inclusive ranges
        self.processed_data = processed_data if processed_data is not None else {}

    def mark(self, file_idx, begin, end):
        from .data_filter import merge_ranges

        key = str(file_idx)
        self.processed_data[key] = [list(r) for r in merge_ranges(
            [tuple(r) for r in self.processed_data.get(key, [])] + [(begin, end)])]

    def is_processed(self, file_idx, record_no):
        from .data_filter import is_processed

        return is_processed(record_no, self.processed_data.get(str(file_idx), []))

    def merge(self, other):
        """Union with another pod's consumed ranges (``other``: DataCheckpoint or its ``processed_data`` dict) -- what the
        trainers of a new stage exchange after an in-place rescale so that the re-created reader skips everything ANY pod
        has already trained on."""
        data = other.processed_data if isinstance(other, DataCheckpoint) else (other or {})
        for key, ranges in data.items():
            for b, e in ranges:
                self.mark(key, b, e)
        return self


class EpochAttr(Serializable):
    def __init__(self):
        self.epoch_no = None
        self.world_size = None
        self.step_num = None
        self.avg_step_time = None
        self.step_no_of_epoch = None


class TrainStatus(Serializable):
    _nested_dict = {"_epochs": EpochAttr}

    def __init__(self, epoch_no=-1):
        self._epoch_no = epoch_no      # last finished / current epoch
        self.global_step_no = 0
        self._epochs = {}              # str(epoch_no) -> EpochAttr
        self.status = int(edl_train_status.TrainStatus.INITIAL)

    @property
    def epoch_no(self):
        return self._epoch_no

    @epoch_no.setter
    def epoch_no(self, epoch_no):
        assert epoch_no >= 0
        self._epochs.setdefault(str(epoch_no), EpochAttr())
        self._epoch_no = epoch_no

    def next(self):
        """Epoch to resume with (fleet ``TrainStatus.next()`` semantics,
        example/collective/resnet50/train_with_fleet.py:491)."""
        return self._epoch_no + 1

    def get_epoch_attr(self, epoch_no):
        return self._epochs.get(str(epoch_no))

    def update_epoch_attr(self, epoch_no, epoch_attr):
        self._epochs[str(epoch_no)] = epoch_attr

    def get_current_epoch_attr(self):
        return self.get_epoch_attr(self._epoch_no)

    def update_current_epoch_attr(self, epoch_attr):
        return self.update_epoch_attr(self._epoch_no, epoch_attr)


class State(Serializable):
    _nested = {"_data_checkpoint": DataCheckpoint, "_train_status": TrainStatus}

    def __init__(self, total_batch_size, user_defined=None):
        self._default = {"total_batch_size": total_batch_size}
        self._user_defined = user_defined
        self._adjust_func = []
        self._name = unique_name.generate("_edl_state_")
        self._model_path = None
        self._data_checkpoint = DataCheckpoint()
        self._train_status = TrainStatus()

    # ---- (de)serialisation: user_defined is an opaque SerializableBase, callbacks are not stored
    def to_dict(self, filter_names=None):
        return {
            "_default": self._default,
            "_user_defined": self._user_defined.to_json() if self._user_defined is not None else None,
            "_name": self._name,
            "_model_path": self._model_path,
            "_data_checkpoint": self._data_checkpoint.to_dict(),
            "_train_status": self._train_status.to_dict(),
        }

    def from_dict(self, d):
        self._default = d["_default"]
        if self._user_defined is not None and d.get("_user_defined") is not None:
            self._user_defined.from_json(d["_user_defined"])
        self._name = d["_name"]
        self._model_path = d["_model_path"]
        self._data_checkpoint = DataCheckpoint().from_dict(d["_data_checkpoint"])
        self._train_status = TrainStatus().from_dict(d["_train_status"])
        return self

    # ---- interface
    def register_adjust_function(self, f):
        """``f(state, old_world_size, new_world_size)`` (a list of callables is accepted, like the
        reference's ``state.register_adjust_function([adjust])``, tests/unittests/test_train.py:54-65)."""
        if isinstance(f, (list, tuple)):
            self._adjust_func.extend(f)
        else:
            self._adjust_func.append(f)

    def adjust(self, old_world_size, new_world_size):
        for f in self._adjust_func:
            f(self, old_world_size, new_world_size)

    @property
    def name(self): return self._name
    @property
    def model_path(self): return self._model_path
    @model_path.setter
    def model_path(self, p): self._model_path = p
    @property
    def data_checkpoint(self): return self._data_checkpoint
    @property
    def train_status(self): return self._train_status
    @property
    def user_defined(self): return self._user_defined
    @property
    def epoch_no(self): return self._train_status.epoch_no
    @property
    def global_step_no(self): return self._train_status.global_step_no

    @property
    def step_no_of_epoch(self):
        a = self._train_status.get_current_epoch_attr()
        return a.step_no_of_epoch if a is not None else None

    @property
    def total_batch_size(self): return self._default["total_batch_size"]
    @total_batch_size.setter
    def total_batch_size(self, size): self._default["total_batch_size"] = size

    # ---- bookkeeping helpers used by notify_end_one_batch / notify_end_one_epoch
    def end_one_batch(self, world_size, step_time=None):
        ts = self._train_status
        ts.global_step_no = int(ts.global_step_no or 0) + 1
        if ts.epoch_no is None or ts.epoch_no < 0:
            ts.epoch_no = 0
        a = ts.get_current_epoch_attr() or EpochAttr()
        a.epoch_no, a.world_size = ts.epoch_no, world_size
        a.step_no_of_epoch = int(a.step_no_of_epoch or 0) + 1
        if step_time is not None:
            n = a.step_no_of_epoch
            a.avg_step_time = step_time if a.avg_step_time is None else a.avg_step_time + (step_time - a.avg_step_time) / n
        ts.update_current_epoch_attr(a)
        ts.status = int(edl_train_status.TrainStatus.RUNNING)

    def end_one_epoch(self):
        ts = self._train_status
        a = ts.get_current_epoch_attr()
        if a is not None:
            a.step_num = a.step_no_of_epoch
        ts.epoch_no = (ts.epoch_no if ts.epoch_no is not None else -1) + 1
        self._data_checkpoint.processed_data = {}


@handle_errors_until_timeout
def load_from_etcd(etcd, state_name, user_defined=None, timeout=60):
    value = etcd.get_value(constants.ETCD_STATE, state_name)
    if value is None:
        raise EdlTableError("no state record {}".format(etcd.get_full_path(constants.ETCD_STATE, state_name)))
    state = State(total_batch_size=None, user_defined=user_defined)
    state.from_json(bytes_to_string(value))
    return state


@handle_errors_until_timeout
def save_to_etcd(etcd, pod_id, state, timeout=60):
    """Leader-guarded write: succeeds only while ``rank/0 == pod_id``."""
    ok = etcd.txn_put_if_value(constants.ETCD_POD_RANK, constants.ETCD_POD_LEADER, pod_id,
                               [(constants.ETCD_STATE, state.name, state.to_json())])
    if not ok:
        raise EdlEtcdIOError("pod_id:{} is not the leader; state not saved".format(pod_id))


class TorchState(State):
    """State bound to live training objects (the reference's ``PaddleState(exe, program, optimizer)``,
    state.py:203-217).  ``trainer`` may be an ``edl_b200.trainer.StudentTrainer``; ``model`` /
    ``optimizer`` are used for plain PyTorch loops."""

    def __init__(self, total_batch_size=None, user_defined=None, model=None, optimizer=None, trainer=None,
                 batch=0, epoch=0):
        super().__init__(total_batch_size=total_batch_size, user_defined=user_defined)
        self._model, self._optimizer, self._trainer = model, optimizer, trainer
        if epoch:
            self._train_status.epoch_no = epoch
        self._train_status.global_step_no = batch

    def to_dict(self, filter_names=None):
        return super().to_dict(filter_names)

    def tensors(self):
        """Everything that has to go into the model checkpoint."""
        if self._trainer is not None:
            return self._trainer.state_dict()
        out = {}
        if self._model is not None:
            out["model"] = self._model.state_dict()
        if self._optimizer is not None:
            out["optim"] = self._optimizer.state_dict()
        return out

    def load_tensors(self, sd):
        if self._trainer is not None:
            self._trainer.load_state_dict(sd)
            return
        if self._model is not None and "model" in sd:
            self._model.load_state_dict(sd["model"])
        if self._optimizer is not None and "optim" in sd:
            self._optimizer.load_state_dict(sd["optim"])


PaddleState = TorchState  # API-compatible alias used by ported user scripts


# ------------------------------------------------------------------------------------------------
# hyper-parameter rescale policies (doc/edl_collective_design_doc.md:14-17: "keep total batch" vs
# "linear scale"); register them with State.register_adjust_function

def linear_scale_lr(get_lr, set_lr):
    """LR <- LR * new_world / old_world (per-trainer batch fixed, total batch scales)."""
    def adjust(state, old_world, new_world):
        if old_world and new_world and old_world != new_world:
            set_lr(get_lr() * float(new_world) / float(old_world))
            if state.total_batch_size:
                state.total_batch_size = int(state.total_batch_size * new_world / old_world)
    return adjust


def keep_total_batch(set_batch_per_trainer):
    """Per-trainer batch <- total_batch / new_world (total batch and LR fixed)."""
    def adjust(state, old_world, new_world):
        if new_world and state.total_batch_size:
            set_batch_per_trainer(max(1, int(state.total_batch_size) // int(new_world)))
    return adjust
