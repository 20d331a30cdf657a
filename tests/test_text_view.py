"""Debug views: text / text_full output equals the reference's renderer
(fixture recorded from gobblet.py:299-429 by oracle/make_golden.py)."""
import numpy as np
import pytest

from oracle import oracle as O


def test_text_formatting_matches_reference_output(golden):
    from gobblet_rl_b200 import text_view
    d = golden("render_text")
    sq, agent = np.zeros(27, np.int8), 0
    for t, a in enumerate(d["actions"]):
        sq = O.play_turn(sq, agent, int(a))
        agent = 1 - agent
        name = ("player_1", "player_2")[agent]
        assert text_view.render_text(sq.astype(np.float64), t + 1, name, int(a), full=False) == str(d["text"][t])
        assert text_view.render_text(sq.astype(np.float64), t + 1, name, int(a), full=True) == str(d["text_full"][t])
    assert text_view.flatboard_from_squares(sq).tolist() == _flat(sq)


def _flat(sq):
    import ctypes as C
    out = (C.c_int * 9)()
    O.lib().gbo_flatboard(np.ascontiguousarray(sq, np.int8).ctypes.data_as(C.c_void_p), out)
    return list(out)


@pytest.mark.gpu
@pytest.mark.parametrize("mode", ["text", "text_full"])
def test_raw_env_text_render_modes(golden, capsys, mode):
    from gobblet_rl_b200 import gobblet_v1
    d = golden("render_text")
    env = gobblet_v1.raw_env(render_mode=mode)
    env.reset()
    capsys.readouterr()
    for t, a in enumerate(d["actions"]):
        env.step(int(a))
        assert capsys.readouterr().out == str(d[mode][t])
    assert env.board.get_flatboard().tolist() == _flat(env.board.squares.astype(np.int8))
