"""Mirror of the HoloGAN slice of the reference's `core.lightning_module` (BaseGAN :35-102, HOLOGAN :209-237): the
`_target_` of `model.lm` in conf/expt/hologan.yaml, so that `instantiate(cfg.model.lm, cfg, logging_dir)`
(run_network.py:41) builds the B200 path.

With pytorch_lightning installed the classes ARE LightningModules (same hooks: `training_step(batch, batch_idx,
optimizer_idx)`, `configure_optimizers` with the `frequency` schedule, `validation_step`); without it (this image,
SURVEY R6) they fall back to a minimal stand-in base with the attributes the steps use (`log`, `device`), which is what
the tests and `HologanTrainer` drive.  The other GANs of the reference's module (DCGAN, WGAN, R1, piGAN, AniGAN) are out
of scope.  Data loading (`train_dataloader` ...) instantiates `cfg.dataset.*` exactly like the
reference when that node exists.
"""
from __future__ import annotations

import torch
from torch import nn

from .. import ops
from ..config import instantiate

try:                                                    # pragma: no cover - not installed in the build image
    import pytorch_lightning as pl
    _Base = pl.LightningModule
    HAVE_LIGHTNING = True
except ImportError:
    HAVE_LIGHTNING = False

    class _Base(nn.Module):
        """The few LightningModule facilities the HoloGAN steps rely on."""

        def __init__(self):
            super().__init__()
            self.logged = {}
            self.current_epoch = 0
            self.logger = None

        @property
        def device(self):
            return next(self.parameters()).device

        def log(self, name, value, **_kw):
            self.logged[name] = value.detach() if isinstance(value, torch.Tensor) else value


class BaseGAN(_Base):
    def __init__(self, cfg, logging_dir):
        super().__init__()
        self.discriminator = instantiate(cfg.discriminator)
        self.generator = instantiate(cfg.generator)
        self.cfg = cfg
        self.logging_dir = logging_dir
        try:                                            # torchvision only matters for real-image data loading
            from torchvision import transforms
            n = cfg.train.channels_img
            self.transform = transforms.Compose([
                transforms.Resize((cfg.train.img_size, cfg.train.img_size)), transforms.ToTensor(),
                transforms.Normalize(mean=[cfg.train.data_mean] * n, std=[cfg.train.data_std] * n)])
        except ImportError:
            self.transform = None
        self.criterion = instantiate(cfg.train.criterion)
        self.noise_distn = instantiate(cfg.model.noise_distn)
        self.fixed_noise = self.noise_distn.sample((8, cfg.model.noise_dim))

    def training_step(self, batch, batch_idx, optimizer_idx):
        raise NotImplementedError

    def validation_step(self, batch, batch_idx):
        real, _ = batch
        return {"real": real}

    def configure_optimizers(self):
        opt_disc = instantiate(self.cfg.disc_optimiser, self.discriminator.parameters())
        opt_gen = instantiate(self.cfg.gen_optimiser, self.generator.parameters())
        sched_disc = instantiate(self.cfg.optimisation.lr_scheduler, optimizer=opt_disc)
        sched_gen = instantiate(self.cfg.optimisation.lr_scheduler, optimizer=opt_gen)
        return ({"optimizer": opt_disc, "lr_scheduler": sched_disc, "frequency": self.cfg.optimisation.disc_freq},
                {"optimizer": opt_gen, "lr_scheduler": sched_gen, "frequency": self.cfg.optimisation.gen_freq})

    def _loader(self, split):
        from torch.utils.data import DataLoader
        dataset = instantiate(self.cfg.dataset[split], transform=self.transform)
        return DataLoader(dataset, num_workers=self.cfg.train.num_workers, batch_size=self.cfg.train.batch_size)

    def train_dataloader(self):
        return self._loader("train")

    def val_dataloader(self):
        return self._loader("val")

    def test_dataloader(self):
        return self._loader("test")


class HOLOGAN(BaseGAN):
    """`training_step` of the reference (core/lightning_module.py:209-237): BCE-with-logits adversarial loss + latent
    identity loss; optimizer 0 = discriminator, 1 = generator.  On CUDA both losses and their gradients run in the fused
    kernels of `ops.hologan_d_loss / hologan_g_loss`; elsewhere in plain torch with `self.criterion`.  Under bf16
    autocast (Lightning `precision="bf16"`) generator and discriminator run on the B200 kernels."""

    def training_step(self, batch, batch_idx, optimizer_idx):
        real, _ = batch
        z = self.noise_distn.sample((len(real), self.cfg.model.noise_dim)).to(self.device)
        if optimizer_idx == 0:
            with torch.no_grad():                       # the reference detaches fake here (:221): no G graph is needed
                fake = self.generator(z)
            disc_real, _ = self.discriminator(real)
            disc_fake, d_z_pred = self.discriminator(fake)
            if real.is_cuda:
                loss, parts = ops.hologan_d_loss(disc_real, disc_fake, d_z_pred, z)
                self.log("train/d_loss", parts[0])
                self.log("train/q_loss", parts[1])
                return loss
            loss_disc = (self.criterion(disc_real, torch.ones_like(disc_real)) +
                         self.criterion(disc_fake, torch.zeros_like(disc_fake))) / 2
            q_loss = torch.mean((d_z_pred - z) ** 2)
            self.log("train/d_loss", loss_disc)
            self.log("train/q_loss", q_loss)
            return loss_disc + q_loss
        if optimizer_idx == 1:
            fake = self.generator(z)
            output, d_z_pred = self.discriminator(fake)
            if real.is_cuda:
                loss, parts = ops.hologan_g_loss(output, d_z_pred, z)
                self.log("train/g_loss", parts[0])
                self.log("train/q_loss", parts[1])
                return loss
            loss_gen = self.criterion(output, torch.ones_like(output))
            q_loss = torch.mean((d_z_pred - z) ** 2)
            self.log("train/g_loss", loss_gen)
            self.log("train/q_loss", q_loss)
            return loss_gen + q_loss
        raise ValueError("optimizer_idx must be 0 (discriminator) or 1 (generator)")
