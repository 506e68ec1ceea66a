"""Host logic of the backward lanes (engine._Lanes): which events are recorded and waited for.  CUDA streams / events are
replaced by fakes that log the calls, so the hazard bookkeeping is checked without a GPU."""
import contextlib

import pytest
import torch

from vit_ae_plus_plus_b200 import engine


class _Log(list):
    pass


@pytest.fixture
def fake_cuda(monkeypatch):
    log = _Log()
    current = []

    class Stream:
        n = 0

        def __init__(self, device=None, priority=0):
            Stream.n += 1
            self.name, self.priority = f"s{Stream.n}", priority

        def wait_event(self, ev):
            log.append(("wait", self.name, ev.id))

    class Event:
        n = 0

        def __init__(self):
            Event.n += 1
            self.id = Event.n

        def record(self, stream=None):
            log.append(("record", (stream or current[-1]).name, self.id))

    main = Stream()
    main.name = "main"
    current.append(main)

    @contextlib.contextmanager
    def stream_ctx(s):
        current.append(s)
        try:
            yield
        finally:
            current.pop()

    monkeypatch.setattr(torch.cuda, "Stream", Stream)
    monkeypatch.setattr(torch.cuda, "Event", Event)
    monkeypatch.setattr(torch.cuda, "current_stream", lambda *a, **k: current[-1])
    monkeypatch.setattr(torch.cuda, "stream", stream_ctx)
    return log, current


def test_side_lanes_start_after_main_and_publish_their_reads(fake_cuda):
    log, current = fake_cuda
    lanes = engine._Lanes("cuda:0")
    s0, s1 = (s.name for s in lanes.streams)
    ran = []
    lanes.side(lambda: ran.append(current[-1].name), reads=("a", "b"), lane=0)
    lanes.side(lambda: ran.append(current[-1].name), reads=("b",), lane=1)
    assert ran == [s0, s1]                                            # the work was enqueued on the lanes' own streams
    # each call: event on main, the lane waits for it, then a completion event on the lane
    assert log[0][:2] == ("record", "main") and log[1] == ("wait", s0, log[0][2]) and log[2][:2] == ("record", s0)
    assert log[3][:2] == ("record", "main") and log[4] == ("wait", s1, log[3][2]) and log[5][:2] == ("record", s1)
    done0, done1 = log[2][2], log[5][2]
    del log[:]
    lanes.before_write("b")                                           # both lanes read b: main waits for both
    assert sorted(log) == sorted([("wait", "main", done0), ("wait", "main", done1)])
    del log[:]
    lanes.before_write("b")                                           # forgotten once waited for
    assert log == []
    lanes.before_write("a", "never-read")
    assert log == [("wait", "main", done0)]


def test_join_waits_for_dirty_lanes_only_and_clears_the_readers(fake_cuda):
    log, _ = fake_cuda
    lanes = engine._Lanes("cuda:0")
    s0, s1 = (s.name for s in lanes.streams)
    lanes.join()
    assert log == []                                                  # nothing forked: nothing to wait for
    lanes.side(lambda: None, reads=("x",), lane=1)
    del log[:]
    lanes.join()
    assert [e[:2] for e in log] == [("record", s1), ("wait", "main")] and log[1][2] == log[0][2]
    del log[:]
    lanes.before_write("x")                                           # the join ordered everything: no stale reader left
    lanes.join()
    assert log == []


def test_side_lanes_run_below_the_captured_main_lane(fake_cuda, monkeypatch):
    for k in ("VITAE_SIDE_PRIORITY", "VITAE_REDUCE_PRIORITY"):
        monkeypatch.delenv(k, raising=False)
    lanes = engine._Lanes("cuda:0")
    assert [s.priority for s in lanes.streams] == [0, 0]               # main lane graphs are captured at priority -1
