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
        g = self._graphed_step(data1, data2) if train else None
        if g is not None:
            vals = g(data1, data2)
            self.optimizer.step()
            return tuple(vals.tolist())
        if train:
            self.optimizer.zero_grad()
        out = self.model(data1, data2)
        losses = self.loss_functionGVAE2(data1, data2, *out, train=train)
        if train:
            losses[0].backward()
            self.optimizer.step()
        return tuple(torch.stack([l.detach() for l in losses]).tolist())

    def _graphed_step(self, data1, data2):
        """CUDA-graph replay of forward + loss + backward (dvae_b200.graph.GraphedTrainStep) when it has been asked for
        (`self.cuda_graph = True` or DVAE_B200_GRAPH=1), the model trains on one GPU and the batch has the captured shape; the
        first two steps of a shape run eagerly (they create the optimizer state and learn the gradient-buffer layout)."""
        import os
        if not (getattr(self, "cuda_graph", False) or os.environ.get("DVAE_B200_GRAPH", "0") == "1"):
            return None
        if not (self.model.training and data1.is_cuda) or getattr(self.model._engine, "buckets", None) is not None:
            return None
        key = (tuple(data1.shape), tuple(data2.shape))
        st = self.__dict__.setdefault("_graph_state", {})
        ent = st.get(key)
        if ent is None:
            st[key] = 1            # eager steps seen with this shape
            return None
        if isinstance(ent, int):
            if ent < 2:
                st[key] = ent + 1
                return None
            from dvae_b200.graph import GraphedTrainStep
            ent = st[key] = GraphedTrainStep(self, data1, data2)
        return ent

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
            g = self._graphed_step(data1, data2)
            if g is not None:
                vals = g(data1, data2)
                self.optimizer.step()
                account(scalars.push(vals))
                continue
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

    def convert_utterances(self, sources, targets, clamp=True):
        """Many-to-many conversion of whole utterances in one pass (BASELINE config 4; the reference loops over utterances
        in Python, :264-296): `sources` / `targets` are either a [U, 80, T] tensor or a list of U [80, T_i] tensors / arrays
        (any lengths); target i lends its style to source i.  Everything between the raw mels and the finished outputs runs
        on the device: chunking_mel (:335-348: 64-frame chunks, last one zero padded, a whole zero chunk when T % 64 == 0),
        ONE encoder pass over all source + target chunks, per-utterance style means (group-mean kernel), ONE decoder pass
        for reconstruction + conversion, postnet on the converted half, time-concat and clamp (:288-296).

        Returns (recons, converted): [U, 80, n*64] tensors for tensor input, lists of [80, n_i*64] tensors for list input;
        `converted` = decode + postnet residual, clamped to [0, 1] like :296 unless clamp=False."""
        from dvae_b200 import ops
        from dvae_b200.engine import N_MELS, T_FRAMES
        m = self.model
        dev = next(m.parameters()).device
        dt, E = m._dt, m._engine
        S, L = m.speaker_size, m.latent_dim
        as_tensor = torch.is_tensor(sources)

        def flatten(mels):
            """-> (flat fp32 device buffer, lengths)"""
            if torch.is_tensor(mels):
                assert mels.dim() == 3 and mels.shape[1] == N_MELS, "expected [U, 80, T]"
                return mels.to(device=dev, dtype=torch.float32).contiguous().view(-1), [int(mels.shape[2])] * int(mels.shape[0])
            parts = [torch.as_tensor(x).to(device=dev, dtype=torch.float32).contiguous() for x in mels]
            assert all(p.dim() == 2 and p.shape[0] == N_MELS for p in parts), "expected a list of [80, T_i]"
            return torch.cat([p.view(-1) for p in parts]), [int(p.shape[1]) for p in parts]
        src_flat, src_len = flatten(sources)
        trg_flat, trg_len = flatten(targets)
        U = len(src_len)
        assert len(trg_len) == U, "one target utterance per source utterance"

        def plan(lengths):
            """Integer bookkeeping of chunking_mel (host, exact): offsets of every utterance and of its chunks."""
            n = np.asarray([t // T_FRAMES + 1 for t in lengths], dtype=np.int64)
            first = np.concatenate([[0], np.cumsum(n)]).astype(np.int32)
            mel_off = np.concatenate([[0], np.cumsum(np.asarray(lengths, dtype=np.int64) * N_MELS)])[:-1].astype(np.int64)
            out_off = np.concatenate([[0], np.cumsum(n * T_FRAMES * N_MELS)]).astype(np.int64)
            utt = np.repeat(np.arange(len(lengths), dtype=np.int32), n)
            return n, first, mel_off, out_off, utt
        n_s, first_s, off_s, out_off_s, utt_s = plan(src_len)
        n_t, first_t, off_t, _, utt_t = plan(trg_len)
        Ns, Nt = int(first_s[-1]), int(first_t[-1])
        up = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev, non_blocking=True)
        d_first_s, d_off_s, d_out_off_s, d_utt_s = up(first_s), up(off_s), up(out_off_s), up(utt_s)
        d_first_t, d_off_t, d_utt_t = up(first_t), up(off_t), up(utt_t)
        d_len_s, d_len_t = up(np.asarray(src_len, dtype=np.int32)), up(np.asarray(trg_len, dtype=np.int32))
        with torch.no_grad():
            W, P, B = m._prepared(), m._param_dict(), m._buffer_dict()
            x_cl = torch.empty((Ns + Nt, T_FRAMES, N_MELS), device=dev, dtype=ops.act_dtype(dt))
            x_cl[:Ns] = ops.chunk_mel(dt, src_flat, d_off_s, d_len_s, d_first_s, d_utt_s, Ns)
            x_cl[Ns:] = ops.chunk_mel(dt, trg_flat, d_off_t, d_len_t, d_first_t, d_utt_t, Nt)
            heads, _ = E.encode_rows(W, P, B, x_cl, 1, m.training, None)          # [Ns + Nt, 2L] fp32
            s_mu = heads[:Ns, :S].contiguous()
            t_mu = heads[Ns:, :S].contiguous()
            c_mu = heads[:Ns, 2 * S:2 * S + (L - S)]

            def utt_mean(mu, gid_rows):
                acc, cnt = ops.group_accumulate(ops.MODE_MEAN, mu, mu, gid_rows, U)
                out, _ = ops.group_finalize(ops.MODE_MEAN, acc, cnt, d_utt_s, Ns, S, want_b=False)
                return out
            z = torch.cat([torch.cat([utt_mean(s_mu, d_utt_s), c_mu], dim=-1),
                           torch.cat([utt_mean(t_mu, d_utt_t), c_mu], dim=-1)], dim=0).contiguous()
            z_act = torch.empty(z.shape, device=dev, dtype=ops.act_dtype(dt))
            ops.prep_cast(dt, z, z_act)
            rec, rec32 = E.decode_rows(W, P, B, z_act, 1, m.training, None)        # rows [0, Ns): recons, [Ns, 2Ns): converted
            post = E.postnet_rows(W, P, B, rec[Ns:], 1, m.training, None)
            total = int(out_off_s[-1])
            recons = torch.empty((total,), device=dev, dtype=torch.float32)
            converted = torch.empty((total,), device=dev, dtype=torch.float32)
            ops.unchunk_mel(dt, rec32[:Ns], None, recons, d_out_off_s, d_first_s, d_utt_s)
            ops.unchunk_mel(dt, rec32[Ns:], post, converted, d_out_off_s, d_first_s, d_utt_s, clamp=(0.0, 1.0) if clamp else None)
        if as_tensor:
            n0 = int(n_s[0])
            return recons.view(U, N_MELS, n0 * T_FRAMES), converted.view(U, N_MELS, n0 * T_FRAMES)
        split = lambda flat: [flat[int(out_off_s[u]):int(out_off_s[u + 1])].view(N_MELS, int(n_s[u]) * T_FRAMES) for u in range(U)]
        return split(recons), split(converted)

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
        n_conv = min(2, len(source_utt_fp))                      # the reference converts the first two source utterances (:264)
        if n_conv == 0:
            return
        sources = [np.load(source_utt_fp[i]) for i in range(n_conv)]
        targets = [np.load(target_utt_fp[np.random.choice(len(target_utt_fp), 1)[0]]) for _ in range(n_conv)]
        recons, converted = self.convert_utterances(sources, targets)   # chunking, encode, style mean, decode, concat, clamp: on device
        for i in range(n_conv):
            stem = Path(source_utt_fp[i]).stem.split("_")
            utterance_id = stem[-2] if len(stem) >= 2 else stem[-1]
            print('convert utterance: {} from --->{} to --->{}'.format(utterance_id, src_spk, trg_spk))
            recons_voice = recons[i].cpu().numpy()
            converted_voice = converted[i].cpu().numpy()
            n_frames = recons_voice.shape[1]
            source_full = np.zeros((sources[i].shape[0], n_frames), dtype=np.float32)
            source_full[:, :sources[i].shape[1]] = sources[i]
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
