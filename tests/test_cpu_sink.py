"""Video sink wire format — runs without a GPU."""
import numpy as np


def test_frames_through_ffmpeg_sink_wire_format(tmp_path, monkeypatch):
    """FFmpegSink (render.py:58-91 of the reference: rawvideo rgb24 on stdin): a stand-in `ffmpeg` executable records its
    argv and stdin, so the wire format is checked without the real encoder: frame bytes in order, -s WxH, -framerate,
    libx264 / yuv420p / preset, audio mux flags."""
    import json
    import os
    import stat

    from maua_stylegan2_b200.render import FFmpegSink

    fake = tmp_path / "bin"
    fake.mkdir()
    exe = fake / "ffmpeg"
    exe.write_text("#!/usr/bin/env python3\nimport sys, json\n"
                   f"open({str(tmp_path / 'argv.json')!r}, 'w').write(json.dumps(sys.argv[1:]))\n"
                   f"open({str(tmp_path / 'stdin.bin')!r}, 'wb').write(sys.stdin.buffer.read())\n")
    exe.chmod(exe.stat().st_mode | stat.S_IEXEC)
    monkeypatch.setenv("PATH", str(fake) + os.pathsep + os.environ["PATH"])
    rng = np.random.default_rng(0)
    batches = [rng.integers(0, 256, (3, 8, 12, 3), dtype=np.uint8) for _ in range(4)]
    sink = FFmpegSink(str(tmp_path / "out.mp4"), 12, 8, 29.97, audio_file="track.wav", offset=1.5, duration=4.0,
                      preset="slow")
    for b in batches:
        sink(b)
    sink.close()
    argv = json.loads((tmp_path / "argv.json").read_text())
    assert (tmp_path / "stdin.bin").read_bytes() == b"".join(b.tobytes() for b in batches)
    joined = " ".join(argv)
    for want in ("-f rawvideo", "-pix_fmt rgb24", "-framerate 29.97", "-s 12x8", "-i pipe:", "-ss 1.5", "-t 4.0",
                 "-i track.wav", "-vcodec libx264", "-pix_fmt yuv420p", "-preset slow", "-b:a 320K", "-ac 2"):
        assert want in joined, (want, joined)
    assert argv[-1].endswith("out.mp4")
