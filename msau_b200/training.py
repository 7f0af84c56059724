"""The reference's alternative trainer over the CUDA engine: ``UNetLoss``, ``get_optimizer``, ``Trainer``.

Mirrors model/training/cost.py:6-65, model/training/optimizer.py:4-30 and model/training/trainer.py:13-207 (SURVEY.md
section 8(f) row 4): one-hot targets, ``0.5 * CE(logits) + 0.5 * CE(aux_logits)`` over ALL pixels, RMSprop by default, the step
learning-rate schedule ``0.001 * 0.95 ** (epoch // 10)``, checkpoints on the best validation loss or every 8 epochs.

What runs where: the arg-max of the one-hot targets, both cross-entropies, the masked accuracy, the whole backward pass and
the optimiser update are CUDA kernels behind ``include/msau_b200.h`` (``msau_onehot_argmax``, ``msau_loss_backward_ex``,
``msau_optimizer_step``); this module is host plumbing with the reference's names, arguments and return values.

Differences, on purpose:
  * ``optimizer_name is "rmsprop"`` in the reference (optimizer.py:14) compares identity; it is ``==`` here;
  * only ``cost_name="cross_entropy"`` exists in the reference's ``UNetLoss`` in any working form (the other branch returns an
    un-constructed activation class, cost.py:38-40), so that is the one built;
  * ``Trainer.train`` drives the fused ``MSAUWrapper.train_step`` (no autograd), ``UNetLoss(...)(logits, tgt, kwargs)`` keeps the
    autograd-style call (``loss.backward()`` then ``optimizer.step()``) for code that wants the reference's own loop.
"""
from __future__ import annotations

import os
import time
from typing import Dict, Optional

import torch

from . import _lib
from .model import MSAUWrapper

_OWNERS: Dict[int, "MSAUWrapper"] = {}


def _owner_of(logits: torch.Tensor) -> MSAUWrapper:
    """The model whose last training forward produced ``logits`` (UNetLoss gets tensors, not the model)."""
    for m in list(_OWNERS.values()):
        if m._last is not None and m._last[3] == logits.data_ptr():
            return m
    raise _lib.MsauError("UNetLoss: logits do not come from the last training forward of a registered msau_b200.MSAUWrapper "
                         "(call msau_b200.training.register(model) once, or use Trainer)")


def register(model: MSAUWrapper) -> MSAUWrapper:
    _OWNERS[id(model)] = model
    return model


class UNetLoss(torch.nn.Module):
    """model/training/cost.py:6-65.  ``forward(logits, tgt, kwargs) -> (acc, loss, final_loss)``."""

    def __init__(self, kwargs):
        super().__init__()
        self.cost_name = kwargs.get("cost_name", "cross_entropy")
        if self.cost_name != "cross_entropy":
            raise NotImplementedError("msau_b200 implements cost_name='cross_entropy' (the reference's only working cost)")
        self.class_weights = kwargs.get("class_weights", None)

    def forward(self, logits, tgt, kwargs):
        aux_logits = kwargs.get("aux_logits", None)
        aux_tgt = kwargs.get("aux_tgt", None)
        model = _owner_of(logits)
        return model.unet_loss(logits, tgt, aux_logits=aux_logits, aux_tgt=aux_tgt, class_weights=self.class_weights)


class FusedOptimizer:
    """What ``get_optimizer`` returns: ``zero_grad()`` / ``step()`` / ``param_groups`` like a torch optimiser, the update itself is
    ``msau_optimizer_step`` on the model's flat parameter / gradient buffers (no per-tensor foreach)."""

    def __init__(self, model: MSAUWrapper, name: str, lr: float, weight_decay: float, momentum: float = 0.9):
        self.model, self.name = model, name
        self.param_groups = [dict(lr=lr, weight_decay=weight_decay, momentum=momentum)]

    def zero_grad(self, set_to_none: bool = True):
        self.model.zero_grad(set_to_none=set_to_none)

    def step(self):
        g = self.param_groups[0]
        m = self.model
        # p.grad are views of the flat gradient buffer (MSAUWrapper._assign_grads); parameters without a gradient (grad None:
        # the dead last-block attention) hold zeros there, and a zero gradient is a no-op for all three rules at wd = 0
        if self.name == "rmsprop":
            m.optimizer_step("rmsprop", g["lr"], (0.99, 0.0), 1e-8, 0.0, g["weight_decay"])
        elif self.name == "momentum":
            m.optimizer_step("momentum", g["lr"], (g["momentum"], 0.0), 0.0, 0.0, g["weight_decay"])
        else:
            m.optimizer_step("adam", g["lr"], (0.9, 0.999), 1e-8, 0.0, g["weight_decay"])

    def fused_args(self) -> dict:
        """keyword arguments that make ``MSAUWrapper.train_step`` apply this optimiser"""
        g = self.param_groups[0]
        if self.name == "rmsprop":
            return dict(optimizer="rmsprop", lr=g["lr"], betas=(0.99, 0.0), eps=1e-8, max_norm=0.0, weight_decay=g["weight_decay"])
        if self.name == "momentum":
            return dict(optimizer="momentum", lr=g["lr"], betas=(g["momentum"], 0.0), eps=0.0, max_norm=0.0, weight_decay=g["weight_decay"])
        return dict(optimizer="adam", lr=g["lr"], betas=(0.9, 0.999), eps=1e-8, max_norm=0.0, weight_decay=g["weight_decay"])


def get_optimizer(model, kwargs={}):
    """model/training/optimizer.py:4-30: "momentum" -> SGD(momentum 0.9), "rmsprop" (default) -> RMSprop, else Adam; the
    learning rate defaults to 0.001 and ``lr_decay_rate`` is passed as weight decay (:7,12,16,21)."""
    name = kwargs.get("optimizer", "rmsprop")
    lr = kwargs.get("learning_rate", 0.001)
    wd = kwargs.get("lr_decay_rate", 0.0)
    if name == "momentum":
        return FusedOptimizer(model, "momentum", lr, wd, kwargs.get("momentum", 0.9))
    if name == "rmsprop":
        return FusedOptimizer(model, "rmsprop", lr, wd)
    return FusedOptimizer(model, "adam", 0.001 if lr is None else lr, wd)      # torch.optim.Adam's own default lr


class Trainer:
    """model/training/trainer.py:13-207.  ``data_provider.next_data('train' | 'val')`` returns ``(batch_x [B,C,H,W], batch_tgt,
    batch_tgt_aux)`` with one-hot targets [B,n_class,H,W]; ``size_val``, ``batchsize_tr``, ``restart_val_runner()``, ``stop_all()``
    as in the reference's generators."""

    def __init__(self, net: MSAUWrapper, opt_kwargs={}, cost_kwargs={}):
        self.net = register(net)
        self.opt_kwargs = opt_kwargs
        self.use_auxiliary_loss = cost_kwargs.get("use_auxiliary_loss", True)
        self.cost_kwargs = {"aux_logits": None, "aux_tgt": None} if self.use_auxiliary_loss else cost_kwargs
        self.cost_type = cost_kwargs.get("cost_name", "cross_entropy")
        self.criterion = UNetLoss(self.cost_kwargs)
        self.class_weights = cost_kwargs.get("class_weights", None)
        self.history = []

    def _initialize(self, output_path):
        self.optimizer = get_optimizer(self.net, self.opt_kwargs)
        if output_path is not None:
            os.makedirs(os.path.abspath(output_path), exist_ok=True)

    def adjust_lr(self, epoch):
        """trainer.py:45-49"""
        lr = 0.001 * (0.95 ** (epoch // 10))
        for g in self.optimizer.param_groups:
            g["lr"] = lr
        return lr

    def _loss_spec(self):
        return dict(mode=1, weight_main=0.5 if self.use_auxiliary_loss else 1.0, weight_aux=0.5 if self.use_auxiliary_loss else 0.0,
                    class_weights=self.class_weights)

    def train(self, data_provider, output_path, restore_path=None, batch_steps_per_epoch=1024, epochs=250, gpu_device="0",
              max_spat_dim=5000000, use_graph=False):
        save_path = os.path.join(output_path, "model") if output_path is not None else None
        if epochs == 0:
            return save_path
        self._initialize(output_path)
        dev = torch.device("cuda", int(gpu_device)) if str(gpu_device).isdigit() else torch.device("cuda")
        self.net.to(dev)
        val_size = data_provider.size_val
        if restore_path is not None:
            self.net.load_weights(restore_path)
        best = 100000.0
        shown = 0
        for epoch in range(epochs):
            lr = self.adjust_lr(epoch)
            t0 = time.time()
            total, total_final, accs = 0.0, 0.0, []
            self.net.train()
            for _ in range(batch_steps_per_epoch):
                bx, bt, ba = data_provider.next_data("train")
                if bx is None:
                    break
                skipped = 0
                while bx.size()[2] * bx.size()[3] > max_spat_dim:       # trainer.py:114-120
                    bx, bt, ba = data_provider.next_data("train")
                    skipped += 1
                    if skipped > 100:
                        return save_path
                bx = bx.float().to(dev)
                tgt = self.net.onehot_argmax(bt.to(dev))
                aux = self.net.onehot_argmax(ba.to(dev)) if self.use_auxiliary_loss else None
                loss = self.net.train_step(bx, tgt, loss_spec=self._loss_spec(), labels_aux=aux, use_graph=use_graph,
                                           **self.optimizer.fused_args())
                accs.append(self.net.last_accuracy())
                total += float(loss)
                total_final += float(self.net._loss_main)
                shown += bx.size()[0]
            n = max(len(accs), 1)
            rec = dict(epoch=epoch + 1, lr=lr, train_loss=total / n, train_final_loss=total_final / n,
                       train_acc=sum(accs) / n if accs else float("nan"), shown=shown, train_s=time.time() - t0)
            vt, vf, vaccs = 0.0, 0.0, []
            self.net.eval()
            for _ in range(val_size):
                bx, bt, ba = data_provider.next_data("val")
                if bx is None:
                    break
                v = self.evaluate_batch(bx.float().to(dev), bt.to(dev), ba.to(dev))
                vaccs.append(v[0]); vt += v[1]; vf += v[2]
            if val_size != 0:
                rec.update(val_loss=vt / val_size, val_final_loss=vf / val_size, val_acc=sum(vaccs) / max(len(vaccs), 1))
                data_provider.restart_val_runner()
            self.history.append(rec)
            val_total = rec.get("val_loss", 0.0)
            if output_path is not None and (val_total < best or (epoch + 1) % 8 == 0):     # trainer.py:196-202
                best = min(best, val_total)
                self.net.save(save_path + str(epoch + 1))
        data_provider.stop_all()
        return save_path

    def evaluate_batch(self, bx, bt, ba):
        """validation half of trainer.py:156-177: forward + UNetLoss, no parameter update -> (acc, loss, final_loss) floats."""
        was = self.net.training
        self.net.train()                       # the loss kernels need the training workspace; nothing is updated
        with torch.enable_grad():
            _, logits, aux = self.net(bx)
            acc, loss, final = self.net.unet_loss(logits, bt, aux_logits=aux if self.use_auxiliary_loss else None, aux_tgt=ba,
                                                  class_weights=self.class_weights)
        self.net.train(was)
        return acc, float(loss.detach()), float(final) if final is not None else float(loss.detach())
