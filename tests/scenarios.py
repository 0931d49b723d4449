"""The reference's own known-answer tests, re-expressed against a backend object.

Each function restates one TEST_CASE/SECTION of the reference's Catch2 suite
(unit_test/bboard/board_logic.cpp, step_utility_test.cpp) with the same fixture calls and the
same REQUIREs, so the same scenario can be run on
  * the compiled reference (oracle.reference()),
  * the plain-C restatement (oracle.restatement()),
  * the CUDA engine (tests/gpu_backend.py: fixtures built on the host, Step on the GPU).

Backend API (mirrors bboard::State methods): zero_state(), put_agent, put_agents_in_corners, kill,
plant_bomb(set_item), spawn_flame, put_item, set_bomb_direction(logical_index, dir), step(s, moves).
States are numpy arrays of shape (1,) with dtype oracle.STATE_DT.
"""
IDLE, UP, DOWN, LEFT, RIGHT, BOMB = range(6)

PASSAGE, RIGID, WOOD, BOMB_ITEM, FLAMES = 0, 1, 2 << 8, 3, 4 << 16
AGENT0 = 1 << 24
BOMB_LIFETIME = 10
FLAME_LIFETIME = 4
MAX_BOMBS_PER_AGENT = 5


def is_flame(c):
    return (int(c) >> 16) == 4


def board(s, x, y):
    return int(s["board"][0, y, x])


def agent(s, i):
    return s["agents"][0, i]


def bomb(s, logical):
    return int(s["bombs"][0, (int(s["bombs_index"][0]) + logical) % 20])


def require_agent(s, a, x, y):
    """REQUIRE_AGENT, board_logic.cpp:11-17"""
    assert int(agent(s, a)["x"]) == x, (a, agent(s, a), x, y)
    assert int(agent(s, a)["y"]) == y, (a, agent(s, a), x, y)
    assert board(s, x, y) == AGENT0 + a, (a, board(s, x, y))


def several_steps(be, times, s, m):
    for _ in range(times):
        be.step(s, m)


def place_bombs_horizontally(be, s, ag, bombs):
    """board_logic.cpp:34-46"""
    m = [IDLE] * 4
    for _ in range(bombs):
        m[ag] = BOMB
        be.step(s, m)
        m[ag] = RIGHT
        be.step(s, m)


# ---------------------------------------------------------------- [step function]
def basic_movement(be):                      # board_logic.cpp:55-83
    s = be.zero_state()
    be.put_agents_in_corners(s, 0, 1, 2, 3)
    m = [IDLE] * 4
    m[0] = RIGHT; be.step(s, m); require_agent(s, 0, 1, 0)
    m[0] = DOWN; be.step(s, m); require_agent(s, 0, 1, 1)
    m[0] = LEFT; be.step(s, m); require_agent(s, 0, 0, 1)
    m[0] = UP; be.step(s, m); require_agent(s, 0, 0, 0)
    m[3] = UP; be.step(s, m); require_agent(s, 3, 0, 9)


def obstacle_collision(be):                  # :85-102
    s = be.zero_state()
    be.put_agents_in_corners(s, 0, 1, 2, 3)
    m = [IDLE] * 4
    be.put_item(s, 1, 0, RIGID)
    m[0] = RIGHT; be.step(s, m); require_agent(s, 0, 0, 0)
    m[0] = DOWN; be.step(s, m); require_agent(s, 0, 0, 1)


def movement_against_flames(be):             # :104-119
    s = be.zero_state()
    m = [IDLE] * 4
    be.put_agents_in_corners(s, 0, 1, 2, 3)
    be.spawn_flame(s, 1, 1, 2)
    m[0] = RIGHT
    be.step(s, m)
    assert agent(s, 0)["dead"]
    assert board(s, 0, 0) == PASSAGE


def _dest_collision_fixture(be):             # :121-133
    s = be.zero_state()
    be.put_agent(s, 0, 1, 0)
    be.put_agent(s, 2, 1, 1)
    be.kill(s, 2, 3)
    return s, [IDLE] * 4


def dest_collision_two(be):                  # :134-143
    s, m = _dest_collision_fixture(be)
    m[0] = RIGHT; m[1] = LEFT
    be.step(s, m)
    require_agent(s, 0, 0, 1); require_agent(s, 1, 2, 1)


def dest_collision_dead(be):                 # :144-153
    s, m = _dest_collision_fixture(be)
    m[0] = RIGHT; m[1] = LEFT
    be.kill(s, 1)
    be.step(s, m)
    require_agent(s, 0, 1, 1)


def dest_collision_four(be):                 # :154-170
    s, m = _dest_collision_fixture(be)
    be.put_agent(s, 1, 0, 2)
    be.put_agent(s, 1, 2, 3)
    m[:] = [RIGHT, LEFT, DOWN, UP]
    be.step(s, m)
    require_agent(s, 0, 0, 1); require_agent(s, 1, 2, 1)
    require_agent(s, 2, 1, 0); require_agent(s, 3, 1, 2)


def chain_against_obstacle(be):              # :181-197
    s = be.zero_state()
    for i in range(4):
        be.put_agent(s, i, 0, i)
    be.put_item(s, 4, 0, RIGID)
    be.step(s, [RIGHT] * 4)
    for i in range(4):
        require_agent(s, i, i, 0)


def two_on_one(be):                          # :198-220 (executes defect D1 in the reference)
    s = be.zero_state()
    be.put_agent(s, 0, 0, 0); be.put_agent(s, 2, 0, 1)
    be.put_agent(s, 1, 0, 2); be.put_agent(s, 1, 1, 3)
    be.step(s, [RIGHT, LEFT, DOWN, DOWN])
    require_agent(s, 0, 0, 0); require_agent(s, 1, 2, 0)
    require_agent(s, 2, 1, 1); require_agent(s, 3, 1, 2)


def move_ouroboros(be):                      # :221-238
    s = be.zero_state()
    be.put_agent(s, 0, 0, 0); be.put_agent(s, 1, 0, 1)
    be.put_agent(s, 1, 1, 2); be.put_agent(s, 0, 1, 3)
    be.step(s, [RIGHT, DOWN, LEFT, UP])
    require_agent(s, 3, 0, 0); require_agent(s, 0, 1, 0)
    require_agent(s, 1, 1, 1); require_agent(s, 2, 0, 1)


def bomb_laying(be):                         # :247-257
    s = be.zero_state()
    m = [IDLE] * 4
    be.put_agents_in_corners(s, 0, 1, 2, 3)
    m[0] = BOMB; be.step(s, m)
    assert board(s, 0, 0) == AGENT0
    m[0] = DOWN; be.step(s, m)
    assert board(s, 0, 0) == BOMB_ITEM


def bomb_block_simple(be):                   # :258-266
    s = be.zero_state()
    m = [IDLE] * 4
    be.put_agents_in_corners(s, 0, 1, 2, 3)
    be.plant_bomb(s, 1, 0, 0)
    m[0] = RIGHT; be.step(s, m)
    require_agent(s, 0, 0, 0)


def bomb_block_complex(be):                  # :267-285
    s = be.zero_state()
    for i in range(4):
        be.put_agent(s, i, 0, i)
    be.step(s, [RIGHT, RIGHT, RIGHT, BOMB])
    require_agent(s, 0, 0, 0); require_agent(s, 1, 1, 0); require_agent(s, 2, 2, 0)
    be.step(s, [IDLE, IDLE, IDLE, RIGHT])
    require_agent(s, 3, 4, 0)


def bomb_ouroboros_block(be):                # :286-306
    s = be.zero_state()
    be.put_agent(s, 0, 0, 0); be.put_agent(s, 1, 0, 1)
    be.put_agent(s, 1, 1, 2); be.put_agent(s, 0, 1, 3)
    be.step(s, [BOMB] * 4)
    be.step(s, [RIGHT, DOWN, LEFT, UP])
    require_agent(s, 0, 0, 0); require_agent(s, 1, 1, 0)
    require_agent(s, 2, 1, 1); require_agent(s, 3, 0, 1)


def _explosion_fixture(be):                  # :310-317
    s = be.zero_state()
    be.kill(s, 2, 3)
    be.put_agent(s, 5, 5, 0)
    return s, [IDLE] * 4


def bomb_goes_off(be):                       # :319-330
    s, m = _explosion_fixture(be)
    m[0] = BOMB; be.step(s, m)
    m[0] = UP; several_steps(be, BOMB_LIFETIME - 1, s, m)
    assert board(s, 5, 5) == BOMB_ITEM
    be.step(s, m)
    assert is_flame(board(s, 5, 5))


def destroy_objects_and_agents(be):          # :331-345
    s, m = _explosion_fixture(be)
    be.put_item(s, 6, 5, WOOD)
    be.put_agent(s, 4, 5, 1)
    m[0] = BOMB; be.step(s, m)
    m[0] = UP; several_steps(be, BOMB_LIFETIME, s, m)
    assert agent(s, 1)["dead"]
    assert is_flame(board(s, 4, 5)) and is_flame(board(s, 6, 5))


def keep_rigid(be):                          # :346-357
    s, m = _explosion_fixture(be)
    be.put_item(s, 6, 5, RIGID)
    m[0] = BOMB; be.step(s, m)
    m[0] = UP; several_steps(be, BOMB_LIFETIME, s, m)
    assert board(s, 6, 5) == RIGID


def kill_only_one_wood(be):                  # :358-369
    s, m = _explosion_fixture(be)
    be.put_item(s, 7, 5, WOOD); be.put_item(s, 8, 5, WOOD)
    s["agents"][0, 0]["bombStrength"] = 5
    be.plant_bomb(s, 6, 5, 0, True)
    several_steps(be, BOMB_LIFETIME, s, m)
    assert is_flame(board(s, 7, 5)) and not is_flame(board(s, 8, 5))


def max_bomb_limit(be):                      # :370-381
    s, m = _explosion_fixture(be)
    s["agents"][0, 0]["maxBombCount"] = 2
    assert int(agent(s, 0)["bombCount"]) == 0
    place_bombs_horizontally(be, s, 0, 4)
    assert board(s, 5, 5) == BOMB_ITEM and board(s, 6, 5) == BOMB_ITEM and board(s, 7, 5) == PASSAGE
    assert int(agent(s, 0)["bombCount"]) == 2


def _flame_fixture(be):                      # :384-389
    s = be.zero_state()
    be.put_agents_in_corners(s, 0, 1, 2, 3)
    return s, [IDLE] * 4


def flame_lifetime(be):                      # :392-401
    s, m = _flame_fixture(be)
    be.spawn_flame(s, 5, 5, 4)
    be.step(s, m)
    several_steps(be, FLAME_LIFETIME - 2, s, m)
    assert is_flame(board(s, 5, 5))
    be.step(s, m)
    assert not is_flame(board(s, 5, 5))


def flame_vanish_completely(be):             # :402-414
    s, m = _flame_fixture(be)
    be.spawn_flame(s, 5, 5, 4)
    be.step(s, m)
    for i in range(5):
        assert is_flame(board(s, 5 + i, 5)) and is_flame(board(s, 5 - i, 5))
        assert is_flame(board(s, 5, 5 + i)) and is_flame(board(s, 5, 5 - i))


def flame_only_vanish_own(be):               # :415-426
    s, m = _flame_fixture(be)
    be.spawn_flame(s, 5, 5, 4)
    be.step(s, m)
    be.spawn_flame(s, 6, 6, 4)
    several_steps(be, FLAME_LIFETIME - 1, s, m)
    assert is_flame(board(s, 6, 5)) and is_flame(board(s, 5, 6))
    assert not is_flame(board(s, 5, 5))


def chained_two_bombs(be):                   # :440-449
    s = be.zero_state()
    m = [IDLE] * 4
    be.put_agents_in_corners(s, 0, 1, 2, 3)
    be.plant_bomb(s, 5, 5, 0, True)
    be.step(s, m)
    be.plant_bomb(s, 4, 5, 1, True)
    several_steps(be, BOMB_LIFETIME - 1, s, m)
    assert int(s["bombs_count"][0]) == 0
    assert is_flame(board(s, 6, 5))


def chained_covered_by_agent(be):            # :450-469
    s = be.zero_state()
    m = [IDLE] * 4
    be.put_agent(s, 5, 5, 0); be.put_agent(s, 4, 5, 1)
    be.kill(s, 2, 3)
    m[0] = BOMB; be.step(s, m)
    m[1] = BOMB; be.step(s, m)
    m[0] = m[1] = DOWN
    several_steps(be, BOMB_LIFETIME - 2, s, m)
    assert int(s["bombs_count"][0]) == 2
    be.step(s, m)
    assert int(s["bombs_count"][0]) == 0
    assert int(s["flames_count"][0]) == 2


def _kick_fixture(be):                       # :474-484
    s = be.zero_state()
    m = [IDLE] * 4
    be.put_agent(s, 0, 1, 0)
    s["agents"][0, 0]["canKick"] = 1
    be.plant_bomb(s, 1, 1, 0, True)
    s["agents"][0, 0]["maxBombCount"] = MAX_BOMBS_PER_AGENT
    m[0] = RIGHT
    return s, m


def kick_one_agent_one_bomb(be):             # :486-500
    s, m = _kick_fixture(be)
    be.kill(s, 1, 2, 3)
    be.step(s, m)
    require_agent(s, 0, 1, 1)
    assert board(s, 2, 1) == BOMB_ITEM
    for i in range(4):
        assert board(s, 2 + i, 1) == BOMB_ITEM
        be.step(s, m)
        m[0] = IDLE


def kick_against_flame(be):                  # :501-515 (literal Item::FLAMES without queue entry)
    s, m = _kick_fixture(be)
    be.kill(s, 1, 2, 3)
    be.put_item(s, 5, 1, FLAMES)
    be.step(s, m)
    m[0] = IDLE
    several_steps(be, 3, s, m)
    assert is_flame(board(s, 5, 1))
    assert int(s["bombs_count"][0]) == 0
    assert int(s["flames_count"][0]) == 1
    f = s["flames"][0, int(s["flames_index"][0]) % 20]
    assert (int(f["x"]), int(f["y"])) == (5, 1)


def kick_bomb_bomb_collision(be):            # :516-531
    s, m = _kick_fixture(be)
    be.kill(s, 1, 2, 3)
    be.plant_bomb(s, 7, 7, 0, True)
    be.set_bomb_direction(s, 1, UP)
    for _ in range(6):
        be.step(s, m)
        m[0] = IDLE
    assert (bomb(s, 0) & 0xF) == 6
    assert (bomb(s, 1) & 0xF) == 7 and ((bomb(s, 1) >> 4) & 0xF) == 2


def kick_bomb_bomb_static(be):               # :533-548
    s, m = _kick_fixture(be)
    be.kill(s, 1, 2, 3)
    be.plant_bomb(s, 7, 6, 0, True)
    be.put_item(s, 7, 0, WOOD)
    be.set_bomb_direction(s, 1, UP)
    for _ in range(7):
        be.step(s, m)
        m[0] = IDLE
    assert (bomb(s, 0) & 0xF) == 6
    assert (bomb(s, 1) & 0xF) == 7 and ((bomb(s, 1) >> 4) & 0xF) == 1


def bounce_back_agent(be):                   # :549-562
    s, m = _kick_fixture(be)
    be.kill(s, 2, 3)
    be.put_agent(s, 0, 2, 1)
    m[1] = UP
    be.plant_bomb(s, 2, 2, 0, True)
    be.set_bomb_direction(s, 1, UP)
    be.step(s, m)
    require_agent(s, 0, 0, 1); require_agent(s, 1, 0, 2)
    assert (bomb(s, 0) & 0xF) == 1 and (bomb(s, 1) & 0xF) == 2


def bounce_back_complex_chain(be):           # :563-580
    s, m = _kick_fixture(be)
    be.kill(s, 2, 3)
    be.put_agent(s, 0, 2, 1)
    m[1] = UP
    be.plant_bomb(s, 2, 2, 0, True)
    be.plant_bomb(s, 0, 3, 0, True)
    be.set_bomb_direction(s, 1, UP)
    be.set_bomb_direction(s, 2, UP)
    be.step(s, m)
    require_agent(s, 0, 0, 1); require_agent(s, 1, 0, 2)
    assert board(s, 0, 3) == BOMB_ITEM and board(s, 1, 1) == BOMB_ITEM and board(s, 2, 2) == BOMB_ITEM


def bounce_back_super_complex(be):           # :581-600 (no REQUIREs in the reference; state trace is the pin)
    s, m = _kick_fixture(be)
    be.kill(s, 3)
    be.put_agent(s, 0, 2, 1)
    be.put_agent(s, 1, 3, 2)
    be.put_item(s, 2, 1, RIGID)
    m[1] = UP
    m[2] = BOMB
    be.plant_bomb(s, 0, 3, 0, True)
    be.set_bomb_direction(s, 1, UP)
    for _ in range(3):
        be.step(s, m)
        m[0] = m[1] = IDLE
        m[2] = LEFT


def bounce_back_wall(be):                    # :602-614
    s, m = _kick_fixture(be)
    be.kill(s, 1, 3)
    be.put_agent(s, 1, 3, 2)
    be.put_item(s, 2, 1, RIGID)
    m[2] = LEFT
    s["agents"][0, 2]["canKick"] = 1
    be.plant_bomb(s, 0, 3, 0, True)
    be.step(s, m)
    require_agent(s, 2, 1, 3)
    assert board(s, 0, 3) == BOMB_ITEM


def stepping_on_bombs(be):                   # :615-634
    s, m = _kick_fixture(be)
    be.put_agent(s, 6, 3, 0); be.put_agent(s, 6, 4, 1); be.put_agent(s, 6, 5, 2)
    m[0] = m[1] = m[2] = IDLE
    be.plant_bomb(s, 5, 6, 3, True)
    be.plant_bomb(s, 6, 6, 2, True)
    be.put_agent(s, 6, 6, 3)
    m[3] = IDLE; be.step(s, m); require_agent(s, 3, 6, 6)
    m[3] = LEFT; be.step(s, m); require_agent(s, 3, 6, 6)


STEP_SCENARIOS = [
    basic_movement, obstacle_collision, movement_against_flames,
    dest_collision_two, dest_collision_dead, dest_collision_four,
    chain_against_obstacle, two_on_one, move_ouroboros,
    bomb_laying, bomb_block_simple, bomb_block_complex, bomb_ouroboros_block,
    bomb_goes_off, destroy_objects_and_agents, keep_rigid, kill_only_one_wood, max_bomb_limit,
    flame_lifetime, flame_vanish_completely, flame_only_vanish_own,
    chained_two_bombs, chained_covered_by_agent,
    kick_one_agent_one_bomb, kick_against_flame, kick_bomb_bomb_collision, kick_bomb_bomb_static,
    bounce_back_agent, bounce_back_complex_chain, bounce_back_super_complex, bounce_back_wall,
    stepping_on_bombs,
]


# ---------------------------------------------------------------- [step utilities]
def util_dest_pos(be):                       # step_utility_test.cpp:38-61
    s = be.zero_state()
    for i in range(4):
        be.put_agent(s, i, 0, i)
    d = be.fill_dest_pos(s, [DOWN, LEFT, RIGHT, UP])
    assert list(d) == [0, 1, 0, 0, 3, 0, 3, -1]


def util_fix_switch(be):                     # :63-84
    s = be.zero_state()
    for i in range(4):
        be.put_agent(s, i, 0, i)
    d = be.fill_dest_pos(s, [RIGHT, RIGHT, LEFT, LEFT])
    d = be.fix_switch_move(s, d)
    assert list(d) == [1, 0, 1, 0, 2, 0, 2, 0]


def _roots(be, places, moves, kill=()):
    s = be.zero_state()
    for i, (x, y) in enumerate(places):
        be.put_agent(s, x, y, i)
    be.kill(s, *kill)
    d = be.fill_dest_pos(s, moves)
    return be.resolve_dependencies(s, d)


def util_dependencies(be):                   # :86-173
    n, dep, roots = _roots(be, [(0, 0), (1, 0), (8, 4), (9, 8)], [RIGHT, RIGHT, RIGHT, IDLE])
    assert 1 in list(roots)
    n, dep, roots = _roots(be, [(0, 0), (1, 0), (8, 8), (9, 8)], [RIGHT, RIGHT, RIGHT, IDLE])
    assert 1 in list(roots) and 3 in list(roots)
    n, dep, roots = _roots(be, [(0, 0), (1, 0), (2, 0), (3, 0)], [RIGHT] * 4)
    assert 3 in list(roots)
    n, dep, roots = _roots(be, [(0, 0), (1, 0), (1, 1), (0, 1)], [RIGHT, DOWN, LEFT, UP])
    assert roots[0] == -1 and n == 0
    n, dep, roots = _roots(be, [(0, 0), (1, 0), (1, 1), (0, 1)], [RIGHT, DOWN, LEFT, UP], kill=(1,))
    assert 0 in list(roots) and 1 in list(roots)


UTIL_SCENARIOS = [util_dest_pos, util_fix_switch, util_dependencies]


class Recorder:
    """Wraps a backend and records every Step as (state_before, moves, state_after)."""

    def __init__(self, be):
        self.be = be
        self.transitions = []

    def __getattr__(self, name):
        return getattr(self.be, name)

    def step(self, s, m):
        before = s.copy()
        r = self.be.step(s, m)
        self.transitions.append((before, [int(x) for x in m], s.copy()))
        return r
