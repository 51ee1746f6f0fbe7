"""Loader half of the reference's streaming runtime (bin/stream.py:23-77), same class and method
names so that ``utils/audiodec.py``-style subclasses work unchanged.

``AudioCodec`` is the abstract loader: subclasses provide ``_load_encoder`` / ``_load_decoder``
(the two plug points of the reference, bin/stream.py:38-45); ``load_transmitter`` /
``load_receiver`` then move the objects to their device and warm the causal state with
``receptive_length`` zeros, exactly like the reference (bin/stream.py:56-77).

``AudioCodecStreamer`` is the duplex real-time streamer (bin/stream.py:80-366): encoder thread,
decoder thread, three queues, latency / frame-drop accounting.  The audio-device part needs
``sounddevice`` (not in this image); ``process_frames`` feeds it synthetic frames through the same
``_process`` path so it can be exercised without audio hardware."""
from __future__ import annotations

import abc
import os
import queue
import threading
import time

import numpy as np
import torch
import yaml


class AudioCodec(abc.ABC):
    def __init__(self, tx_device: str = "cpu", rx_device: str = "cpu", receptive_length: int = 8192):
        self.tx_device, self.rx_device = tx_device, rx_device
        self.receptive_length = receptive_length
        self.tx_encoder = self.rx_encoder = self.decoder = None

    @abc.abstractmethod
    def _load_encoder(self, checkpoint):
        ...

    @abc.abstractmethod
    def _load_decoder(self, checkpoint):
        ...

    def _load_config(self, checkpoint, config_name="config.yml"):
        # the config lives next to the checkpoint file (bin/stream.py:48-53)
        with open(os.path.join(os.path.dirname(checkpoint), config_name)) as f:
            return yaml.load(f, Loader=yaml.Loader)

    def load_transmitter(self, encoder_checkpoint):
        assert os.path.exists(encoder_checkpoint), f"{encoder_checkpoint} does not exist!"
        self.tx_encoder = self._load_encoder(encoder_checkpoint)
        self.tx_encoder.eval().to(self.tx_device)
        self.tx_encoder.initial_encoder(self.receptive_length, self.tx_device)
        print("Load tx_encoder: %s" % encoder_checkpoint)

    def load_receiver(self, encoder_checkpoint, decoder_checkpoint):
        assert os.path.exists(encoder_checkpoint), f"{encoder_checkpoint} does not exist!"
        self.rx_encoder = self._load_encoder(encoder_checkpoint)
        self.rx_encoder.eval().to(self.rx_device)
        zq = self.rx_encoder.initial_encoder(self.receptive_length, self.rx_device)
        print("Load rx_encoder: %s" % encoder_checkpoint)
        assert os.path.exists(decoder_checkpoint), f"{decoder_checkpoint} does not exist!"
        self.decoder = self._load_decoder(decoder_checkpoint)
        self.decoder.eval().to(self.rx_device)
        self.decoder.initial_decoder(zq)
        print("Load decoder: %s" % decoder_checkpoint)


class AudioCodecStreamer(abc.ABC):
    """Microphone -> encoder thread -> decoder thread -> speakers (bin/stream.py:80-366)."""

    def __init__(self, input_device, output_device, input_channels: int = 1, output_channels: int = 1,
                 frame_size: int = 512, sample_rate: int = 48000, gain: float = 1.0, max_latency: float = 0.1,
                 tx_encoder=None, tx_device: str = "cpu", rx_encoder=None, decoder=None, rx_device: str = "cpu"):
        self.input_device, self.output_device = input_device, output_device
        self.input_channels, self.output_channels = input_channels, output_channels
        self.frame_size, self.sample_rate = frame_size, sample_rate
        self.gain, self.max_latency = gain, max_latency
        self.tx_encoder, self.tx_device = tx_encoder, tx_device
        self.rx_encoder, self.decoder, self.rx_device = rx_encoder, decoder, rx_device
        print(f"Encoder device: {tx_device}")
        print(f"Decoder device: {rx_device}")
        self.encoder_queue, self.decoder_queue, self.output_queue = queue.Queue(), queue.Queue(), queue.Queue()
        self.latency_queue = queue.Queue()
        self.input_dump, self.output_dump = [], []
        self.input_dump_filename = self.output_dump_filename = None
        self.frame_drops = self.n_frames = 0
        self.encoder_times, self.decoder_times, self.latencies = [], [], []
        self._threads_started = False

    @abc.abstractmethod
    def _encode(self, x):
        ...

    @abc.abstractmethod
    def _decode(self, x):
        ...

    # -- worker threads (bin/stream.py:212-239) -----------------------------------------------------
    def _worker(self, src, dst, device, fn, times, enabled):
        while threading.main_thread().is_alive():
            try:
                x = src.get(timeout=1)
            except queue.Empty:
                continue
            t0 = time.time()
            x = x.to(device)
            with torch.no_grad():
                if enabled():
                    x = fn(x)
            if x.is_cuda:
                torch.cuda.synchronize(x.device)
            times.append(time.time() - t0)
            dst.put(x)

    def _run_encoder(self):
        self._worker(self.encoder_queue, self.decoder_queue, self.tx_device, self._encode, self.encoder_times,
                     lambda: self.tx_encoder is not None)

    def _run_decoder(self):
        self._worker(self.decoder_queue, self.output_queue, self.rx_device, self._decode, self.decoder_times,
                     lambda: self.rx_encoder is not None and self.decoder is not None)

    def _start_threads(self):
        if not self._threads_started:
            threading.Thread(target=self._run_encoder, daemon=True).start()
            threading.Thread(target=self._run_decoder, daemon=True).start()
            self._threads_started = True

    # -- per-frame callback body (bin/stream.py:242-278) ------------------------------------------------
    def _process(self, data):
        frame = torch.from_numpy(data * self.gain).transpose(1, 0).contiguous()     # (channels, frame_size)
        if self.input_dump_filename is not None:
            self.input_dump.append(frame)
        self.encoder_queue.put(frame.unsqueeze(0))
        self.latency_queue.put(time.time())
        try:
            out = self.output_queue.get_nowait()
            latency = time.time() - self.latency_queue.get_nowait()
            self.latencies.append(latency)
            if latency > self.max_latency:          # too late: flush everything, count the dropped frames
                for q in (self.encoder_queue, self.decoder_queue, self.output_queue):
                    with q.mutex:
                        q.queue.clear()
                while not self.latency_queue.empty():
                    self.frame_drops += 1
                    self.latency_queue.get_nowait()
        except queue.Empty:
            out = torch.zeros(1, self.output_channels, self.frame_size)
        out = out.squeeze(0).detach().cpu()
        self.n_frames += 1
        if self.output_dump_filename is not None:
            self.output_dump.append(out)
        return out.transpose(1, 0).contiguous().numpy()

    def _callback(self, indata, outdata, frames, _time, status):
        if status:
            print(status)
        outdata[:] = self._process(indata)

    def process_frames(self, frames, realtime=False):
        """Drive the streamer with an iterable of (frame_size, channels) float32 numpy frames instead of a
        sound card; returns the list of output frames (zeros until the pipeline has filled)."""
        self._start_threads()
        outs = []
        period = self.frame_size / self.sample_rate
        for f in frames:
            t0 = time.time()
            outs.append(self._process(np.asarray(f, dtype=np.float32)))
            if realtime:
                time.sleep(max(0.0, period - (time.time() - t0)))
        return outs

    def statistics(self):
        ms = lambda v: (float(np.mean(v) * 1e3), float(np.std(v) * 1e3)) if len(v) else (float("nan"), float("nan"))
        return {"encoder_ms": ms(self.encoder_times), "decoder_ms": ms(self.decoder_times),
                "latency_ms": ms(self.latencies), "frame_drops": self.frame_drops, "n_frames": self.n_frames}

    def _exit(self):
        s = self.statistics()
        print("#" * 80)
        print("encoder processing time (ms):      %.2f +- %.2f" % s["encoder_ms"])
        print("decoder processing time (ms):      %.2f +- %.2f" % s["decoder_ms"])
        print("system latency (ms):               %.2f +- %.2f" % s["latency_ms"])
        print("frame drops:                       %d (%.2f%%)" % (self.frame_drops, 100.0 * self.frame_drops / max(1, self.n_frames)))
        print("#" * 80)
        self._write_dumps()

    def _write_dumps(self):
        # bin/stream.py:285-293: concatenate the frames seen by the callback, clamp to [-1, 1], write PCM16 wavs
        from ..wavio import write_wav_pcm16
        for name, frames in ((self.input_dump_filename, self.input_dump), (self.output_dump_filename, self.output_dump)):
            if name is None or not frames:
                continue
            wav = torch.clamp(torch.cat(frames, dim=-1), -1.0, 1.0)           # (channels, n_samples)
            write_wav_pcm16(name, wav.transpose(1, 0).numpy(), self.sample_rate)
            print("Wrote %s (%d samples)" % (name, wav.shape[-1]))
        self.input_dump, self.output_dump = [], []

    def enable_filedump(self, input_stream_file: str = None, output_stream_file: str = None):
        if input_stream_file is None and output_stream_file is None:
            raise Exception("At least one of input_stream_file and output_stream_file must be specified.")
        fix = lambda n: n if n is None or n.endswith(".wav") else n + ".wav"
        self.input_dump_filename, self.output_dump_filename = fix(input_stream_file), fix(output_stream_file)

    def run(self, latency):
        self._start_threads()
        try:
            import sounddevice as sd          # lazy, like bin/stream.py:350
            with sd.Stream(device=(self.input_device, self.output_device), samplerate=self.sample_rate,
                           blocksize=self.frame_size, dtype=np.float32, latency=latency,
                           channels=(self.input_channels, self.output_channels), callback=self._callback):
                print("### starting stream [press Return to quit] ###")
                input()
                self._exit()
        except KeyboardInterrupt:
            self._exit()
        except Exception as e:
            print(type(e).__name__ + ": " + str(e))
