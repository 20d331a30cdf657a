"""Debug views from the engine state: the `text` / `text_full` render modes
of the reference (gobblet.py:299-429) rebuilt from `gbl_export_squares` output.  Pure formatting; no rules."""
import numpy as np

_BLANK = " " * 7 + "|" + " " * 7 + "|" + " " * 7
_RULE = "_" * 7 + "|" + "_" * 7 + "|" + "_" * 7


def flatboard_from_squares(squares):
    """Signed piece number of the visible piece per square (view of board.py:159-177)."""
    sq = np.asarray(squares).reshape(3, 9)
    top = np.abs(sq).argmax(axis=0)
    return sq[top, np.arange(9)]


def _size_symbol(v):          # top-of-stack view: signed SIZE 1..3
    if v == 0:
        return "- "
    return f"+{int((v + 1) // 2)}" if v > 0 else f"{int(v // 2)}"


def _piece_symbol(v):         # full view: signed piece number 1..6
    if v == 0:
        return "- "
    return f"+{int(v)}" if v > 0 else f"{int(v)}"


def _row(cells, r):           # cells are column-major: 0 3 6 / 1 4 7 / 2 5 8
    return f"  {cells[r]}   |   {cells[r + 3]}  |   {cells[r + 6]}  "


def _grid(cell_sets):
    """One 3x3 grid per entry of cell_sets, side by side."""
    join = lambda parts: "  ".join(parts)  # noqa: E731
    k = len(cell_sets)
    lines = []
    for r in range(3):
        lines.append(join([_BLANK] * k))
        lines.append(join([_row(c, r) for c in cell_sets]))
        lines.append(join([_RULE if r < 2 else _BLANK] * k))
    return lines


def render_text(squares, turn, agent_selection, action, full=False):
    pos, piece = action % 9, action // 9 + 1
    sq = np.asarray(squares)
    if full:
        head = f"TURN: {turn}, AGENT: {agent_selection}, ACTION: {action}, POSITION: {pos}, PIECE: {piece}"
        title = " " * 9 + "SMALL" + " " * 9 + "  " + " " * 10 + "MED" + " " * 10 + "  " + " " * 9 + "LARGE" + " " * 9 + "  "
        cells = [[_piece_symbol(v) for v in sq[9 * lvl: 9 * lvl + 9]] for lvl in range(3)]
        lines = [head, title] + _grid(cells)
    else:
        head = f"TURN: {turn}, AGENT: {agent_selection}, ACTION: {action}, POSITION: {pos}, PIECE: {(piece + 1) // 2}"
        lines = [head] + _grid([[_size_symbol(v) for v in flatboard_from_squares(sq)]])
    return "\n".join(lines) + "\n\n"
