"""Learning-rate schedule of the reference training loop (host side, pure Python like the reference's).

Mirrors `WarmupMultiBatchScheduler` (lib/train_utils/lr_scheduler.py:10-66) and the way `get_optim_params`
(lib/train_utils/utils.py:12-43) builds it from the yml (configs/faster/sniper_res101_e2e.yml:104-111: lr 0.015,
lr_step '5.33', warmup from 0.0005 over 1000 updates, factor 0.1).  `__call__(num_update)` has the reference's
semantics: num_update is the optimizer's update count AFTER the increment (mxnet/optimizer.py `_update_count`
precedes `_get_lr`), i.e. the first update of a run asks for scheduler(1).
"""


class WarmupMultiBatchScheduler(object):
    def __init__(self, step, factor=1, warmup=False, warmup_lr=0, warmup_step=0, base_lr=0.01):
        if not (isinstance(step, list) and len(step) >= 1):
            raise AssertionError("step must be a non-empty list")
        for i, _step in enumerate(step):
            if i != 0 and step[i] <= step[i - 1]:
                raise ValueError("Schedule step must be an increasing integer list")
            if _step < 1:
                raise ValueError("Schedule step must be greater or equal than 1 round")
        if factor > 1.0:
            raise ValueError("Factor must be no more than 1 to make lr reduce")
        self.base_lr = base_lr          # mxnet.lr_scheduler.LRScheduler.base_lr (set by the optimizer to learning_rate)
        self.step = step
        self.cur_step_ind = 0
        self.factor = factor
        self.count = 0
        self.warmup = warmup
        self.warmup_lr = warmup_lr
        self.warmup_step = warmup_step

    def __call__(self, num_update):
        if self.warmup and num_update < self.warmup_step:
            return self.warmup_lr + num_update * (self.base_lr - self.warmup_lr) / self.warmup_step
        while self.cur_step_ind <= len(self.step) - 1:
            if num_update > self.step[self.cur_step_ind]:
                self.count = self.step[self.cur_step_ind]
                self.cur_step_ind += 1
                self.base_lr *= self.factor
            else:
                return self.base_lr
        return self.base_lr


def from_config(lr=0.015, lr_step="5.33", lr_factor=0.1, warmup=True, warmup_lr=0.0005, warmup_step=1000,
                begin_epoch=0, roidb_len=None, batch_size=None, fp16=False, scale=100.0):
    """get_optim_params (lib/train_utils/utils.py:12-43): (scheduler, optimizer kwargs).  roidb_len / batch_size turn
    the epoch boundaries of lr_step into update counts; with roidb_len=None the decay steps are effectively disabled
    (benchmarks / synthetic runs shorter than an epoch).  fp16: the reference divides lr and warmup_lr by TRAIN.scale
    and multiplies wd by it because its losses are scaled by TRAIN.scale (utils.py:22-31)."""
    lr_epoch = [float(e) for e in str(lr_step).split(",")]
    diff = [e - begin_epoch for e in lr_epoch if e > begin_epoch]
    if roidb_len is None or batch_size is None:
        iters = [1 << 60]
    else:
        iters = [int(e * roidb_len / batch_size) for e in diff] or [1 << 60]
    if fp16:
        warmup_lr = warmup_lr / scale
        lr = lr / scale
    return WarmupMultiBatchScheduler(iters, lr_factor, warmup, warmup_lr, warmup_step, base_lr=lr)
