"""Trainer / converter base class with the reference's interface (model/variational_base_vae.py:30-360).

`train.py` builds `ConvolutionalMulVAE` (a subclass) and calls `run_training` / `voice_conversion_mel`; those
entry points, their arguments, the checkpoint naming (`DisentangledVAE_VCTK_{epoch}.pth`, state_dict only) and the
returned tuples are kept.  The numerical work (model forward/backward, loss, conversion encode/decode, per-utterance
style mean) runs on dvae_b200 kernels; plotting, tensorboard and the WaveNet vocoder are optional host-side extras
that are skipped when their third-party packages are absent (they are outside the accelerated path).
"""
from __future__ import annotations

import os
from glob import glob
from pathlib import Path

import numpy as np
import torch

try:  # optional, host-side only
    from tqdm import tqdm
except Exception:  # pragma: no cover
    def tqdm(x, **_):
        return x

device = torch.device("cuda" if torch.cuda.is_available() else "cpu")


class _NullWriter:
    def add_scalar(self, *a, **k):
        pass


def _summary_writer(path):
    try:
        from tensorboardX import SummaryWriter
        return SummaryWriter(path)
    except Exception:
        return _NullWriter()


class VariationalBaseModelVAE():
    def __init__(self, dataset, width, height, channels, latent_sz, learning_rate, device, log_interval, batch_size,
                 normalize=False, flatten=True):
        self.dataset = dataset
        self.width = width
        self.height = height
        self.channels = channels
        self.input_sz = (channels, width, height)
        self.latent_sz = latent_sz
        self.lr = learning_rate
        self.device = device
        self.log_interval = log_interval
        self.normalize_data = normalize
        self.flatten_data = flatten
        self.model = None       # set by subclasses
        self.optimizer = None
        self.batch_size = batch_size

    def loss_function(self):
        raise NotImplementedError

    # ------------------------------------------------------------------ one optimisation step (:58-70)
    def step(self, data1, data2, speaker_ids, train=False):
        """zero_grad -> forward -> loss -> backward -> optimizer; returns the 8 loss terms as Python floats.
        `speaker_ids` is accepted and unused, like the reference (the style group is the (x1, x2) pair: SURVEY F2).
        The 8 scalars come back in ONE device->host copy instead of eight `.item()` syncs."""
        if train:
            self.optimizer.zero_grad()
        out = self.model(data1, data2)
        losses = self.loss_functionGVAE2(data1, data2, *out, train=train)
        if train:
            losses[0].backward()
            self.optimizer.step()
        return tuple(torch.stack([l.detach() for l in losses]).tolist())

    def train(self, train_loader, epoch, logging_func=print):
        """One epoch (:74-101).  Returns the reference's 7-tuple of summed loss terms.

        Same arithmetic and the same return value as the reference loop; the host side is pipelined: batch k+1 is
        copied to the device while batch k computes (`DevicePrefetcher`) and the eight loss scalars of a step are read
        back one step late, without stalling the launch queue (`AsyncScalars`)."""
        from dvae_b200.data import AsyncScalars, DevicePrefetcher
        self.model.train()
        tot = [0.0] * 8
        dev = next(self.model.parameters()).device
        scalars = AsyncScalars(8, dev)
        last_style_kl = 0.0

        def account(vals):
            nonlocal tot, last_style_kl
            if vals is not None:
                tot = [a + b for a, b in zip(tot, vals)]
                last_style_kl = vals[7]
        for batch_idx, (data1, data2, speaker_ids) in enumerate(tqdm(DevicePrefetcher(train_loader, dev))):
            self.optimizer.zero_grad()
            out = self.model(data1, data2)
            losses = self.loss_functionGVAE2(data1, data2, *out, train=True)
            losses[0].backward()
            self.optimizer.step()
            account(scalars.push(torch.stack([l.detach() for l in losses])))
        account(scalars.flush())
        if hasattr(train_loader.dataset, "shuffle_data"):
            train_loader.dataset.shuffle_data()
        logging_func('====> Epoch: {} Average loss: {:.4f}'.format(epoch, tot[0] / len(train_loader.dataset)))
        # (recons1, recons2, recons1_hat, recons2_hat, z1_kl, z2_kl, style_kl of the LAST batch -- as the reference)
        return tot[1], tot[2], tot[3], tot[4], tot[5], tot[6], last_style_kl

    def test(self, test_loader, epoch, logging_func=print):
        """Evaluation pass.  (The reference's `test` (:105-123) calls step() with a wrong arity and cannot run;
        this one evaluates the pair loss without updating anything.)"""
        self.model.eval()
        total, n = 0.0, 0
        dev = next(self.model.parameters()).device
        with torch.no_grad():
            for data1, data2, speaker_ids in test_loader:
                vals = self.step(data1.to(dev).float(), data2.to(dev).float(), speaker_ids.view(-1), train=False)
                total += vals[0]
                n += 1
        name = self.model.__class__.__name__
        logging_func(f'====> Test loss {name}: {total / max(n, 1):.4f}')
        return total / max(len(test_loader.dataset), 1)

    # ------------------------------------------------------------------ checkpoints (:127-149)
    def load_last_model(self, checkpoints_path, logging_func=print):
        name = self.model.__class__.__name__
        found = []
        for f in glob(f'{checkpoints_path}/*.pth'):
            model_name, dataset, epoch = Path(f).stem.split('_')
            found.append((int(epoch), f))
        if not found:
            logging_func(f'Training {name} model from scratch...')
            return 1
        start_epoch, last_checkpoint = max(found, key=lambda item: item[0])
        dev = next(self.model.parameters()).device
        self.model.load_state_dict(torch.load(last_checkpoint, map_location=dev))
        logging_func(f'Loading {name} model from last checkpoint ({start_epoch})...')
        return start_epoch + 1

    def update_(self):
        pass

    def run_training(self, train_loader, test_loader, epochs, report_interval, sample_sz=64, reload_model=True,
                     checkpoints_path='', logs_path='', images_path='', estimation_dir='', logging_func=print,
                     start_epoch=None):
        """Epoch loop with resume, scalar logging and periodic checkpoints (:156-202)."""
        start_epoch = self.load_last_model(checkpoints_path, logging_func) if reload_model else 1
        run_name = "DisentangledVAE_VCTK"
        writer = _summary_writer(f'{logs_path}/{run_name}')
        for epoch in range(start_epoch, start_epoch + epochs):
            print('kl coef: ', self.kl_cof)
            r1, r2, r1h, r2h, k1, k2, ks = self.train(train_loader, epoch, logging_func)
            nb = len(train_loader)
            for label, v in (('recons loss1', r1), ('recons loss2', r2), ('recons loss1 hat', r1h),
                             ('recons loss2 hat', r2h), ('Z1 KL loss', k1), ('Z2 kL loss', k2), ('Z Style KL', ks)):
                print('{} epoch_{}: {}'.format(label, epoch, v / nb))
            writer.add_scalar('Loss\\Reconstruction Loss1', r1 / nb, epoch)
            writer.add_scalar('Loss\\Reconstruction Loss2', r2 / nb, epoch)
            writer.add_scalar('Loss\\Z1 KL Loss', k1 / nb, epoch)
            writer.add_scalar('Loss\\Z2 KL Loss', k2 / nb, epoch)
            writer.add_scalar('Loss\\Z KL Style', ks / nb, epoch)
            if epoch % report_interval == 0:
                for d in (images_path, checkpoints_path):
                    if d and not os.path.exists(d):
                        os.makedirs(d, exist_ok=True)
                with torch.no_grad():
                    torch.save(self.model.state_dict(), f'{checkpoints_path}/{run_name}_{epoch}.pth')
                    self.estimate_trained_model(test_loader, checkpoints_path, estimation_dir)

    def estimate_trained_model(self, test_loader, checkpoints_path, estimation_dir):
        """Reconstruct the first test batch in eval mode and save 5 original / reconstructed mels (:205-239).
        PNG plots need matplotlib + librosa; without them the arrays are saved as .npy."""
        logging_epoch = self.load_last_model(checkpoints_path, logging_func=print)
        self.model.eval()
        if estimation_dir and not os.path.exists(estimation_dir):
            os.makedirs(estimation_dir, exist_ok=True)
        dev = next(self.model.parameters()).device
        with torch.no_grad():
            data1, data2, speaker_ids = next(iter(test_loader))
            data1, data2 = data1.to(dev).float(), data2.to(dev).float()
            out = self.model(data1, data2, train=False)
            recons_x1 = out[2]
            for i in range(min(5, data1.shape[0])):
                stem = os.path.join(estimation_dir, str(logging_epoch))
                _save_mel(recons_x1[i].cpu().numpy(), f'{stem}_recons_mel_{i}', 'reconstructed mel spectrogram')
                _save_mel(data1[i].cpu().numpy(), f'{stem}_original_mel_{i}', 'original mel spectrogram')

    # ------------------------------------------------------------------ conversion (:243-330)
    def convert_chunks(self, source_chunks, source_utt, target_chunks, target_utt, n_utts):
        """Batched core of voice conversion (:277-296) for MANY utterances at once.

        source_chunks [Ns,80,64] with utterance index source_utt [Ns] (int32, values < n_utts); target_chunks
        [Nt,80,64] with target_utt [Nt] naming which source utterance each target chunk lends its style to.
        Style of an utterance = mean of style_mu over its chunks (group-mean kernel).  Returns
        (recons [Ns,80,64], converted [Ns,80,64] = decode + postnet residual, not yet clamped)."""
        from dvae_b200 import ops
        m = self.model
        with torch.no_grad():
            s_mu, _, c_mu, _ = m.encode(source_chunks)
            t_mu, _, _, _ = m.encode(target_chunks)
            S = s_mu.shape[1]

            def utt_mean(mu, gid_rows, gid_out):
                mu = mu.contiguous()
                acc, cnt = ops.group_accumulate(ops.MODE_MEAN, mu, mu, gid_rows, n_utts)
                out, _ = ops.group_finalize(ops.MODE_MEAN, acc, cnt, gid_out, gid_out.numel(), S, want_b=False)
                return out
            src_style = utt_mean(s_mu, source_utt, source_utt)
            trg_style = utt_mean(t_mu, target_utt, source_utt)
            recons = m.decode(torch.cat([src_style, c_mu], dim=-1))
            _, converted = m.decode_with_postnet(torch.cat([trg_style, c_mu], dim=-1))
        return recons, converted

    def voice_conversion_mel(self, ckp_path, generation_dir, src_spk, trg_spk, dataset_fp=''):
        """Convert the first two utterances of `src_spk` to the voice of `trg_spk` (:243-330).  The mel-domain
        part runs here; waveform synthesis needs the external WaveNet vocoder + its checkpoint and is skipped
        (the converted mel is saved as .npy) when they are unavailable."""
        save_dir = os.path.join(generation_dir, src_spk + '_to_' + trg_spk)
        os.makedirs(save_dir, exist_ok=True)
        self.load_last_model(ckp_path, logging_func=print)
        self.model.eval()
        dev = next(self.model.parameters()).device
        vocoder = _try_build_vocoder(dev)
        source_utt_fp = np.sort(glob(os.path.join(dataset_fp, src_spk, "*.npy")))
        target_utt_fp = glob(os.path.join(dataset_fp, trg_spk, '*.npy'))
        for i in range(min(2, len(source_utt_fp))):
            source_mel = chunking_mel(np.load(source_utt_fp[i])).to(dev).float()
            rnd_trg = np.random.choice(len(target_utt_fp), 1)[0]
            target_mel = chunking_mel(np.load(target_utt_fp[rnd_trg])).to(dev).float()
            stem = Path(source_utt_fp[i]).stem.split("_")
            utterance_id = stem[-2] if len(stem) >= 2 else stem[-1]
            print('convert utterance: {} from --->{} to --->{}'.format(utterance_id, src_spk, trg_spk))
            zeros_s = torch.zeros(source_mel.shape[0], dtype=torch.int32, device=dev)
            zeros_t = torch.zeros(target_mel.shape[0], dtype=torch.int32, device=dev)
            recons, converted = self.convert_chunks(source_mel, zeros_s, target_mel, zeros_t, 1)
            cat_t = lambda m: torch.cat([m[j] for j in range(m.shape[0])], 1)
            recons_voice = cat_t(recons).cpu().numpy()
            converted_voice = torch.clamp(cat_t(converted), min=0, max=1.0).cpu().numpy()
            source_full = cat_t(source_mel).cpu().numpy()
            _save_mel(source_full, os.path.join(save_dir, f'original_{src_spk}_{utterance_id}'), 'original')
            _save_mel(converted_voice, os.path.join(save_dir, f'convert_{src_spk}_{trg_spk}_{utterance_id}'), 'convert')
            _save_mel(recons_voice, os.path.join(save_dir, f'recons_{src_spk}_{utterance_id}'), 'reconstruct')
            if vocoder is not None:
                _vocode(vocoder, converted_voice.T, os.path.join(
                    save_dir, f'convert_{src_spk}_to_{trg_spk}_{utterance_id}.wav'))


# ---------------------------------------------------------------------- helpers
def chunking_mel(melspectrogram):
    """[80, T] -> [T//64 + 1, 80, 64]: non-overlapping 64-frame chunks, last one zero padded -- a whole zero chunk
    when T % 64 == 0 (model/variational_base_vae.py:335-348)."""
    return _chunk(melspectrogram, 64)


def chunking_mcc(mcc, length=128):
    """Same chunking for WORLD mel-cepstra (:350-360)."""
    return _chunk(mcc, length)


def _chunk(arr, length):
    arr = np.asarray(arr)
    n = arr.shape[1] // length + 1
    out = np.zeros((n, arr.shape[0], length), dtype=arr.dtype)
    for i in range(n):
        piece = arr[:, i * length:(i + 1) * length]
        out[i, :, :piece.shape[1]] = piece
    return torch.from_numpy(out)


def _save_mel(mel, stem, title):
    try:
        import matplotlib
        matplotlib.use("Agg")
        import matplotlib.pyplot as plt
        plt.figure()
        plt.title(title)
        plt.imshow(mel, origin="lower", aspect="auto")
        plt.colorbar(format='%f')
        plt.savefig(stem + '.png')
        plt.close()
    except Exception:
        np.save(stem + '.npy', mel)


def _try_build_vocoder(dev):
    try:
        from preprocessing.processing import build_model
        model = build_model().to(dev)
        ckpt = torch.load('checkpoint_step001000000_ema.pth', map_location=dev)
        model.load_state_dict(ckpt['state_dict'])
        return model
    except Exception as e:  # vocoder is outside the accelerated path
        print(f'[dvae_b200] vocoder unavailable ({type(e).__name__}); saving mels only')
        return None


def _vocode(vocoder, mel_t, path):
    try:
        import soundfile as sf
        from preprocessing.processing import wavegen
        sf.write(path, wavegen(vocoder, mel_t), 16000)
    except Exception as e:
        print(f'[dvae_b200] waveform synthesis skipped ({type(e).__name__})')
