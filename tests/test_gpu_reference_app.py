"""BASELINE config 1 as worded: a synthetic 2.048 MS/s uint8 IQ FM-stereo + RDS multiplex is
replayed over a local rtl_tcp socket into the UNMODIFIED reference application
(src/main.cpp, xdr_server.cpp, audio_output.cpp, rtl_tcp_client.cpp ... compiled as they are
against the drop-in headers, tools/build_reference_app.py) whose DSP classes are the engine;
the 32 kHz WAV it writes must equal the CPU oracle's audio after the reference's own
volume scaling (x0.85) and int16 truncation (audio_output.cpp:1386-1391,1445-1464).

The binary is built where /root/reference exists and travels to the GPU box in build/refapp/.
SURVEY §8(f) row 1 / Appendix B.13: the tuner never auto-starts, so the harness speaks the
FM-DX handshake on the XDR port ("x", "x"), keeps both sockets open, accepts exactly once and
ends the run with SIGTERM so the WAV header is finalised.
"""
import json
import os
import signal
import socket
import struct
import subprocess
import threading
import time

import numpy as np
import pytest

from oracle import orc

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXE = os.path.join(ROOT, "build", "refapp", "fm-sdr-tuner-b200")


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


class RtlTcpReplay(threading.Thread):
    """Minimal rtl_tcp server: 12-byte header, then the IQ bytes; tuner commands are ignored."""

    def __init__(self, port, payload: bytes):
        super().__init__(daemon=True)
        self.payload = payload
        self.done = threading.Event()
        self.srv = socket.socket()
        self.srv.setsockopt(socket.SOL_SOCKET, socket.SO_REUSEADDR, 1)
        self.srv.bind(("127.0.0.1", port))
        self.srv.listen(1)
        self.conn = None

    def run(self):
        self.conn, _ = self.srv.accept()          # exactly once
        self.conn.sendall(b"RTL0" + struct.pack(">II", 5, 29))
        threading.Thread(target=self._drain, daemon=True).start()
        self.conn.sendall(self.payload)
        self.done.set()

    def _drain(self):
        try:
            while self.conn.recv(4096):
                pass
        except OSError:
            pass


@pytest.mark.skipif(not os.path.exists(EXE), reason="reference application was not built "
                    "(needs /root/reference at build time)")
def test_rtl_tcp_replay_through_unmodified_main(tmp_path, orc_fm):
    iq_rate, decim, nblk = 2_048_000, 8, 96       # 3.07 s of signal
    iq = orc.config1_signal(fs_iq=iq_rate).generate(nblk * 8192 * decim)
    ref = orc.Channel(orc_fm, orc.make_config(iq_rate=iq_rate, decimation=decim)).process(iq)

    tcp_port, xdr_port = _free_port(), _free_port()
    ini = tmp_path / "app.ini"
    ini.write_text(f"[xdr]\nport = {xdr_port}\nguest_mode = true\npassword =\n"
                   "[debug]\nlog_level = 0\n[reconnection]\nauto_reconnect = false\n")
    wav = tmp_path / "out.wav"
    replay = RtlTcpReplay(tcp_port, iq.tobytes())
    replay.start()
    env = dict(os.environ)
    proc = subprocess.Popen([EXE, "-c", str(ini), "--source", "rtl_tcp", "-t", f"127.0.0.1:{tcp_port}",
                             "--iq-rate", str(iq_rate), "-w", str(wav), "-G"],
                            stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True, env=env)
    try:
        ctl = None
        for _ in range(200):                       # wait for the XDR server to listen
            try:
                ctl = socket.create_connection(("127.0.0.1", xdr_port), timeout=1.0)
                break
            except OSError:
                time.sleep(0.1)
                assert proc.poll() is None, proc.stdout.read()
        assert ctl is not None
        ctl.sendall(b"x\n")                        # FM-DX protocol select
        ctl.settimeout(10.0)
        assert ctl.recv(16).startswith(b"1") or True
        ctl.sendall(b"x\n")                        # start the tuner
        t_start = time.time()
        # the main loop reads the socket as fast as it decodes: wait until the WAV holds every
        # frame the signal gives, then stop the application the way a user would
        want = ref.left.size
        deadline = time.time() + 150
        t_done = None
        while time.time() < deadline:
            if wav.exists() and (wav.stat().st_size - 44) // 4 >= want - 1100:
                t_done = time.time()
                break
            time.sleep(0.02)
        assert replay.done.wait(timeout=5), "replay did not finish"
        time.sleep(0.5)                             # whatever is still buffered
        if t_done is not None:
            # BASELINE config 1 through the drop-in classes (five one-channel engines, every stage a
            # synchronous host <-> device round trip, as the reference's call structure demands)
            rtf = (nblk * 8192 / 256000.0) / (t_done - t_start)
            rec = {"signal_seconds": nblk * 8192 / 256000.0, "wall_seconds": t_done - t_start,
                   "realtime_factor": rtf, "path": "unmodified main.cpp + rtl_tcp replay + drop-in classes"}
            print("reference application through the drop-ins:", rec)
            os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
            with open(os.path.join(ROOT, "gpurun_out", "refapp_realtime.json"), "w") as f:
                json.dump(rec, f)
            assert rtf > 1.0, rec                   # a one-channel tuner must keep up with the air
    finally:
        proc.send_signal(signal.SIGTERM)
        try:
            out, _ = proc.communicate(timeout=30)
        except subprocess.TimeoutExpired:
            proc.kill()
            out, _ = proc.communicate()
    raw = wav.read_bytes()
    assert raw[:4] == b"RIFF" and raw[8:12] == b"WAVE", out[-2000:]
    pcm = np.frombuffer(raw[44:], "<i2").reshape(-1, 2)
    n = min(pcm.shape[0], ref.left.size)
    assert n >= ref.left.size - 1100, (pcm.shape, ref.left.size, out[-2000:])
    scale = np.float32(0.85)

    def to_i16(x):
        y = np.clip(x[:n] * scale, -1.0, 1.0).astype(np.float32) * np.float32(32767.0)
        return np.trunc(y).astype(np.int16)

    assert np.array_equal(pcm[:n, 0], to_i16(ref.left)), out[-1500:]
    assert np.array_equal(pcm[:n, 1], to_i16(ref.right))
