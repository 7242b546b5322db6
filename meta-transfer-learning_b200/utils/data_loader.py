"""Input pipeline with the reference's class names and batch contracts (utils/data_loader.py).

  SpectrogramParser.parse_audio   wav -> log(1 + |STFT|) with a 20 ms symmetric Hamming window, 10 ms hop,
                                  centred frames with reflect padding, utterance-level mean / std normalisation
                                  (data_loader.py:65-96; the reference delegates the STFT to librosa)
  SpectrogramDataset.sample       K-shot sampler: ((inputs (k,1,F,Tmax), input_sizes (k) int32 raw frames,
                                  input_percentages (k), targets (k,Lmax) int64 PAD=0, target_sizes (k) int32),
                                  (same for the validation shots))           (data_loader.py:245-321)
  AudioDataLoader                 DataLoader whose collate returns (inputs, targets, input_percentages,
                                  input_sizes, target_sizes) sorted by length     (data_loader.py:401-477)
  BucketingSampler                fixed-size index bins                        (data_loader.py:480-500)

Host-side by default (numpy STFT).  With ``audio_conf['device'] = 'cuda'`` (or MTL_FEATURES_ON_DEVICE=1) the K-shot sampler
computes the features on the GPU instead (csrc/spectrogram.cu through mtl_spectrogram): the raw waves go up through one
pinned staging buffer per batch, every utterance's log-spectrogram is written straight into its zero-padded slice of
the (k,1,F,Tmax) batch tensor, and ``sample`` returns CUDA tensors that the trainer's static input slots take with a
device-to-device copy -- the sampling thread of the trainer (transient_trainer.py:127-139) then overlaps disk reads and
uploads with the running meta-step."""
import os
import random

import numpy as np
import pandas as pd
import scipy.signal.windows
import torch
from torch.utils.data import DataLoader, Dataset
from torch.utils.data.sampler import Sampler

from utils.audio import load_audio, load_randomly_augmented_audio

windows = {'hamming': scipy.signal.windows.hamming, 'hann': scipy.signal.windows.hann,
           'blackman': scipy.signal.windows.blackman, 'bartlett': scipy.signal.windows.bartlett}


def stft_magnitude(y, n_fft, hop_length, window):
    """|STFT| (n_fft//2+1, 1 + len(y)//hop) with centred, reflect-padded frames: the librosa.stft defaults the
    reference relied on (data_loader.py:84-86)."""
    y = np.asarray(y, dtype=np.float32)
    pad = n_fft // 2
    if len(y) <= pad:
        raise ValueError("utterance shorter than half a window")
    yp = np.pad(y, pad, mode="reflect")
    n_frames = 1 + (len(yp) - n_fft) // hop_length
    idx = np.arange(n_fft)[None, :] + hop_length * np.arange(n_frames)[:, None]
    frames = yp[idx] * np.asarray(window, dtype=np.float32)[None, :]
    return np.abs(np.fft.rfft(frames, n=n_fft, axis=1)).T.astype(np.float32)


def device_batch_features(waves, n_fft, hop, window, normalize, max_frames, device):
    """k waves (1-D float arrays) -> (inputs (k,1,F,Tmax) float32 on `device`, zero padded; frames per utterance).  Frames
    beyond `max_frames` are dropped AFTER the utterance-level normalisation, like ``parse_audio(...)[:, :src_max_len]``."""
    import ctypes as C
    from mtl_b200 import lib as L
    lib = L.get_lib()
    dev = torch.device(device)
    k = len(waves)
    frames = [1 + len(w) // hop for w in waves]
    t_full = max(frames)
    F = n_fft // 2 + 1
    total = sum(len(w) for w in waves)
    stage = torch.empty(total, dtype=torch.float32).pin_memory()           # one upload for the whole batch
    off, offs = 0, []
    for w in waves:
        stage[off:off + len(w)] = torch.from_numpy(np.ascontiguousarray(w, dtype=np.float32))
        offs.append(off)
        off += len(w)
    with torch.cuda.device(dev):
        st = torch.cuda.current_stream()
        wav_d = stage.to(dev, non_blocking=True)
        win_d = torch.from_numpy(np.asarray(window, dtype=np.float32)).to(dev, non_blocking=True)
        out = torch.zeros(k, 1, F, t_full, dtype=torch.float32, device=dev)
        stat = torch.zeros(2, dtype=torch.float64, device=dev)
        for i, w in enumerate(waves):
            L.check(lib.mtl_spectrogram(C.c_void_p(wav_d.data_ptr() + 4 * offs[i]), len(w), n_fft, hop,
                                        C.c_void_p(win_d.data_ptr()), C.c_void_p(out[i, 0].data_ptr()), t_full,
                                        int(bool(normalize)), C.c_void_p(stat.data_ptr()), C.c_void_p(st.cuda_stream)))
        st.synchronize()                                                   # the staging buffer may be released now
    if t_full > max_frames:
        out = out[:, :, :, :max_frames].contiguous()
        frames = [min(f, max_frames) for f in frames]
    return out, frames


class AudioParser(object):
    def parse_transcript(self, transcript_path):
        raise NotImplementedError

    def parse_audio(self, audio_path):
        raise NotImplementedError


class SpectrogramParser(AudioParser):
    def __init__(self, audio_conf, normalize=False, augment=False):
        super().__init__()
        self.window_stride = audio_conf['window_stride']
        self.window_size = audio_conf['window_size']
        self.sample_rate = audio_conf['sample_rate']
        self.window = windows.get(audio_conf['window'], windows['hamming'])
        self.normalize = normalize
        self.augment = augment
        if audio_conf.get('noise_dir') is not None:
            raise NotImplementedError("noise injection needs sox (data_loader.py:367-399): not available in this build")
        self.noiseInjector = None
        self.noise_prob = audio_conf.get('noise_prob')
        dev = audio_conf.get('device') or ("cuda" if os.environ.get("MTL_FEATURES_ON_DEVICE") == "1" else None)
        self.feature_device = dev if (dev and torch.cuda.is_available()) else None

    def _load_wave(self, audio_path):
        return load_randomly_augmented_audio(audio_path, self.sample_rate) if self.augment else load_audio(audio_path)

    def parse_audio(self, audio_path):
        y = self._load_wave(audio_path)
        n_fft = int(self.sample_rate * self.window_size)
        hop = int(self.sample_rate * self.window_stride)
        spect = torch.from_numpy(np.log1p(stft_magnitude(y, n_fft, hop, self.window(n_fft))))
        if self.normalize:
            spect = (spect - spect.mean()) / spect.std()
        return spect


def _pad_batch(spects, transcripts, pad_id):
    """Zero-pads spectrograms to the longest, PAD-pads label sequences: the tensors both ``sample`` and the
    loader's collate return (field ORDER differs between the two, as in the reference)."""
    k = len(spects)
    t_max = max(s.size(1) for s in spects)
    f_max = max(spects, key=lambda s: s.size(1)).size(0)
    l_max = max(len(t) for t in transcripts)
    inputs = torch.zeros(k, 1, f_max, t_max)
    sizes = torch.zeros(k, dtype=torch.int32)
    pct = torch.zeros(k, dtype=torch.float32)
    targets = torch.full((k, l_max), pad_id).long()
    target_sizes = torch.zeros(k, dtype=torch.int32)
    for i, (s, t) in enumerate(zip(spects, transcripts)):
        n = s.size(1)
        inputs[i, 0, :, :n] = s
        sizes[i] = n
        pct[i] = n / float(t_max)
        target_sizes[i] = len(t)
        if len(t):
            targets[i, :len(t)] = torch.as_tensor(t, dtype=torch.long)
    return inputs, sizes, pct, targets, target_sizes


class SpectrogramDataset(Dataset, SpectrogramParser):
    def __init__(self, vocab, args, audio_conf, manifest_filepath_list, normalize=False, augment=False,
                 input_type="char", is_train=False, partitions=None):
        """Manifests are CSV files of ``wav_path,transcript`` rows; the transcript column is a path to a .txt
        file or the text itself (data_loader.py:171-243,341-361)."""
        self.is_train, self.args, self.vocab = is_train, args, vocab
        self.ids_list = [pd.read_csv(path, header=None).values.tolist() for path in manifest_filepath_list]
        n_manifests = len(manifest_filepath_list)
        self.max_size = max(len(ids) for ids in self.ids_list) * n_manifests
        if is_train and n_manifests > 1:
            self.max_size = 30000
        print("max_size:", self.max_size)
        print("input_type:", input_type)
        self.input_type = input_type
        self.manifest_filepath_list = manifest_filepath_list
        self.proba = []
        self.part_len = self.max_size
        for i, ids in enumerate(self.ids_list):
            if partitions is not None:      # uniform over the leading fraction of each manifest
                self.part_len = max(1, int(len(ids) * partitions[i]))
                p = np.zeros(len(ids))
                p[:self.part_len] = 1 / self.part_len
            else:
                p = np.full(len(ids), 1 / len(ids))
            self.proba.append(p)
        SpectrogramParser.__init__(self, audio_conf, normalize, augment)

    def _load(self, row):
        spect = self.parse_audio(row[0])[:, :self.args.src_max_len]
        return spect, self.parse_transcript(row[1])

    def _device_batch(self, rows):
        """One K-shot batch with the features computed on the GPU: same five fields as ``_pad_batch``, inputs on the device."""
        n_fft, hop = int(self.sample_rate * self.window_size), int(self.sample_rate * self.window_stride)
        waves = [np.asarray(self._load_wave(r[0]), dtype=np.float32) for r in rows]
        transcripts = [self.parse_transcript(r[1]) for r in rows]
        inputs, frames = device_batch_features(waves, n_fft, hop, self.window(n_fft), self.normalize,
                                               self.args.src_max_len, self.feature_device)
        k, t_max = len(rows), inputs.size(3)
        l_max = max(len(t) for t in transcripts)
        sizes = torch.tensor(frames, dtype=torch.int32)
        pct = sizes.float() / float(t_max)
        targets = torch.full((k, l_max), self.vocab.PAD_ID).long()
        target_sizes = torch.tensor([len(t) for t in transcripts], dtype=torch.int32)
        for i, t in enumerate(transcripts):
            if len(t):
                targets[i, :len(t)] = torch.as_tensor(t, dtype=torch.long)
        return inputs, sizes, pct, targets, target_sizes

    def sample(self, k_train, k_val, manifest_id):
        """k_train + k_val rows of one manifest, drawn with replacement (np.random.choice, the reference's RNG)."""
        ids = self.ids_list[manifest_id]
        picks = np.random.choice(np.arange(0, len(ids)), k_train + k_val, p=self.proba[manifest_id], replace=True)
        if self.feature_device is not None:
            return (self._device_batch([ids[i] for i in picks[:k_train]]),
                    self._device_batch([ids[i] for i in picks[k_train:k_train + k_val]]))
        tr = [self._load(ids[i]) for i in picks[:k_train]]
        va = [self._load(ids[i]) for i in picks[k_train:k_train + k_val]]
        pad = self.vocab.PAD_ID
        return (_pad_batch([s for s, _ in tr], [t for _, t in tr], pad),
                _pad_batch([s for s, _ in va], [t for _, t in va], pad))

    def __getitem__(self, index):
        if self.is_train:
            m = len(self.manifest_filepath_list)
            ids = self.ids_list[index % m]
            row = ids[(index // m) % len(ids)]
        else:
            ids = self.ids_list[0]
            row = ids[index % len(ids)]
        return self._load(row)

    def parse_transcript(self, transcript_path):
        if self.input_type != "char":
            raise NotImplementedError("only input_type='char' is supported")
        if transcript_path[-4:] == '.txt':
            with open(transcript_path, 'r', encoding='utf8') as f:
                text = " " + f.read().replace('\n', '').lower()
        else:
            text = transcript_path.replace('\n', '').lower()
        # unknown characters are dropped -- and so is label id 0, exactly like filter(None, ...) in the reference
        return [i for i in (self.vocab.label2id.get(ch) for ch in text) if i]

    def __len__(self):
        return self.part_len

    def uniform_shuffle(self, arr):
        for i in range(32):
            j = random.randint(0, i)
            arr[i], arr[j] = arr[j], arr[i]
        return arr


class LogFBankDataset(SpectrogramDataset):
    def __init__(self, *a, **k):
        raise NotImplementedError("log-filterbank features need python_speech_features (data_loader.py:102-168): "
                                  "not part of this image; use --feat spectrogram")


class AudioDataLoader(DataLoader):
    def __init__(self, pad_token_id, *args, **kwargs):
        super().__init__(*args, **kwargs)
        self.pad_token_id = pad_token_id

        def _collate_fn(batch):
            batch = sorted(batch, key=lambda sample: sample[0].size(1), reverse=True)
            inputs, sizes, pct, targets, target_sizes = _pad_batch([b[0] for b in batch], [b[1] for b in batch],
                                                                   self.pad_token_id)
            return inputs, targets, pct, sizes, target_sizes

        self.collate_fn = _collate_fn


class BucketingSampler(Sampler):
    def __init__(self, data_source, batch_size=1):
        """Consecutive index bins of ``batch_size`` (data assumed sorted by length); bins are shuffled inside."""
        self.data_source = data_source
        ids = list(range(len(data_source)))
        self.bins = [ids[i:i + batch_size] for i in range(0, len(ids), batch_size)]

    def __iter__(self):
        for ids in self.bins:
            np.random.shuffle(ids)
            yield ids

    def __len__(self):
        return len(self.bins)

    def shuffle(self, epoch):
        np.random.shuffle(self.bins)
