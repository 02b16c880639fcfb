// phyx_b200 — solve schedules: which joints may run together, in which order.
//
// Reference stage replaced here:
//   Solver::PrepareIndices   src/Solver.cpp:217-273   greedy grouping of joints into SIMD-wide
//                                                      sets with pairwise distinct bodies
//
// Two schedule families (phyx_b200_schedule):
//   COLOUR      the grouping idea widened from 8 lanes to the whole device: joints are coloured so
//               that no two joints of a colour share a dynamic body; one colour = one level.
//   REPLAY_*    the reference's exact order (PrepareIndices with N = 8 / 4 / 1: groups first, then
//               the scalar tail) turned into dependency levels: a unit (group or single joint) is
//               placed one level after the latest earlier unit that shares a dynamic body with it.
//               Executing levels in order, units of a level in parallel, applies impulses to every
//               body in the same order as the reference's sequential loop.
//
// A level lists its 8-wide units first (each occupying 8 aligned slots, padded with -1 when N = 4)
// and its 1-wide units after them (see Level in common.cuh).
#include "common.cuh"

#include <algorithm>

namespace phyx
{

namespace
{

inline bool is_static(const float4& p) { return p.x == 0.0f && p.y == 0.0f; }   // Solver.cpp:304

// The reference's greedy grouper, restated (Solver.cpp:217-273): repeatedly sweep the remaining
// joints, taking the first N whose bodies are untouched in this round; taken joints are replaced
// by the last remaining one.  Returns groupOffset (multiple of N).
int reference_order(const phyx_contact_joint* joints, int nj, int nb, int N, std::vector<int>& order)
{
    order.resize(nj);
    for (int i = 0; i < nj; ++i) order[i] = i;
    if (N == 1) return 0;
    std::vector<int> pool(order);
    std::vector<int> bodyRound(nb, 0);
    int round = 0, remaining = nj, offset = 0;
    while (remaining >= N)
    {
        int taken = 0;
        ++round;
        for (int i = 0; i < remaining && taken < N;)
        {
            int j = pool[i];
            int b1 = joints[j].body1Index, b2 = joints[j].body2Index;
            if (bodyRound[b1] < round && bodyRound[b2] < round)
            {
                bodyRound[b1] = bodyRound[b2] = round;
                order[offset + taken++] = j;
                pool[i] = pool[--remaining];
            }
            else
                ++i;
        }
        offset += taken;
        if (taken < N) break;
    }
    for (int i = 0; i < remaining; ++i) order[offset + i] = pool[i];
    return offset & ~(N - 1);
}

struct Unit
{
    int first, width, level;   // order[first .. first+width)
};

// How static bodies enter the level assignment of a replay schedule.
enum StaticRule
{
    kStaticsFree = 0,       // ignored: joints on a static body are unordered among themselves (fast schedule)
    kStaticsOrdered = 1,    // joints of one static body sit on non-decreasing levels in sequential order (strict
                            // schedule: exact under every lastIteration state, but levels ratchet along e.g. the
                            // ground's contacts)
    kStaticsSerial = 2      // strictly increasing levels (PHYX_B200_SOLVE_STATIC_DEPS; cross-check only)
};

// reference order -> units (SIMD groups first, then the scalar tail)
void reference_units(const phyx_contact_joint* joints, int nj, int nb, int N, std::vector<int>& order, std::vector<Unit>& units)
{
    int groupOffset = reference_order(joints, nj, nb, N, order);
    units.clear();
    units.reserve(size_t(groupOffset / std::max(N, 1)) + size_t(nj - groupOffset));
    for (int g = 0; N > 1 && g < groupOffset; g += N) units.push_back({ g, N, 0 });
    for (int i = groupOffset; i < nj; ++i) units.push_back({ i, 1, 0 });
}

// Dependency levels + slot layout.  bodyLevel[b] is the earliest level the next unit on body b may take:
// a dynamic body forces a strictly later level (its velocity row is read-modify-written).
void layout_replay(const phyx_contact_joint* joints, int nb, const std::vector<unsigned char>& statics, const std::vector<int>& order,
    std::vector<Unit>& units, StaticRule rule, std::vector<int>& slots, std::vector<int>& slotPos, std::vector<Level>& levels)
{
    std::vector<int> bodyLevel(nb, 0);
    int maxLevel = 0;
    for (Unit& u : units)
    {
        int lvl = 0;
        for (int k = 0; k < u.width; ++k)
        {
            const phyx_contact_joint& j = joints[order[u.first + k]];
            if (rule != kStaticsFree || !statics[j.body1Index]) lvl = std::max(lvl, bodyLevel[j.body1Index]);
            if (rule != kStaticsFree || !statics[j.body2Index]) lvl = std::max(lvl, bodyLevel[j.body2Index]);
        }
        u.level = lvl;   // 0-based level of this unit
        for (int k = 0; k < u.width; ++k)
        {
            const phyx_contact_joint& j = joints[order[u.first + k]];
            bodyLevel[j.body1Index] = lvl + ((rule == kStaticsSerial || !statics[j.body1Index]) ? 1 : 0);
            bodyLevel[j.body2Index] = lvl + ((rule == kStaticsSerial || !statics[j.body2Index]) ? 1 : 0);
        }
        maxLevel = std::max(maxLevel, lvl + 1);
    }
    // bucket by level: wide units first, singles after (stable in unit order)
    std::vector<int> wideCount(maxLevel, 0), singleCount(maxLevel, 0);
    for (const Unit& u : units) (u.width > 1 ? wideCount : singleCount)[u.level]++;
    levels.resize(maxLevel);
    int cursor = 0;
    for (int l = 0; l < maxLevel; ++l)
    {
        levels[l].start = cursor;
        levels[l].grouped_end = cursor + wideCount[l] * 8;
        levels[l].end = levels[l].grouped_end + singleCount[l];
        cursor = (levels[l].end + 7) & ~7;
    }
    slots.assign(cursor, -1);
    slotPos.assign(cursor, 0);   // sequential position of a slot = index of its unit in the reference order
    std::vector<int> wideAt(maxLevel), singleAt(maxLevel);
    for (int l = 0; l < maxLevel; ++l)
    {
        wideAt[l] = levels[l].start;
        singleAt[l] = levels[l].grouped_end;
    }
    for (size_t ui = 0; ui < units.size(); ++ui)
    {
        const Unit& u = units[ui];
        if (u.width > 1)
        {
            for (int k = 0; k < u.width; ++k)
            {
                slots[wideAt[u.level] + k] = order[u.first + k];
                slotPos[wideAt[u.level] + k] = int(ui);
            }
            wideAt[u.level] += 8;
        }
        else
        {
            slotPos[singleAt[u.level]] = int(ui);
            slots[singleAt[u.level]++] = order[u.first];
        }
    }
}

// Greedy colouring in joint order: smallest colour not yet used on either dynamic body.
void build_colours(const phyx_contact_joint* joints, int nj, int nb, const std::vector<unsigned char>& statics, std::vector<int>& slots,
    std::vector<Level>& levels)
{
    int words = 1;
    std::vector<int> colour(nj);
    int numColours = 0;
    for (;;)
    {
        std::vector<uint64_t> used(size_t(nb) * words, 0);
        bool overflow = false;
        numColours = 0;
        for (int j = 0; j < nj && !overflow; ++j)
        {
            int b1 = joints[j].body1Index, b2 = joints[j].body2Index;
            const bool d1 = !statics[b1], d2 = !statics[b2];
            int c = -1;
            for (int w = 0; w < words; ++w)
            {
                uint64_t m = (d1 ? used[size_t(b1) * words + w] : 0) | (d2 ? used[size_t(b2) * words + w] : 0);
                if (~m)
                {
                    c = w * 64 + __builtin_ctzll(~m);
                    break;
                }
            }
            if (c < 0)
            {
                overflow = true;
                break;
            }
            colour[j] = c;
            numColours = std::max(numColours, c + 1);
            if (d1) used[size_t(b1) * words + (c >> 6)] |= uint64_t(1) << (c & 63);
            if (d2) used[size_t(b2) * words + (c >> 6)] |= uint64_t(1) << (c & 63);
        }
        if (!overflow) break;
        words *= 2;
    }
    std::vector<int> count(numColours, 0);
    for (int j = 0; j < nj; ++j) count[colour[j]]++;
    levels.resize(numColours);
    int cursor = 0;
    std::vector<int> at(numColours);
    for (int c = 0; c < numColours; ++c)
    {
        levels[c].start = cursor;
        levels[c].grouped_end = cursor;
        levels[c].end = cursor + count[c];
        at[c] = cursor;
        cursor = (levels[c].end + 31) & ~31;
    }
    slots.assign(cursor, -1);
    for (int j = 0; j < nj; ++j) slots[at[colour[j]]++] = j;
}

} // namespace

int schedule_build(phyx_b200_ctx* c, const phyx_contact_joint* hostJoints, int nj, int mode, int flags)
{
    const int nb = c->bodyCount, ncp = c->contactPointCount;
    const bool deviceColouring = mode == PHYX_B200_SCHEDULE_COLOUR && !(flags & PHYX_B200_SOLVE_HOST_COLOURING);
    // (body1, body2) list: validates indices and (on request) detects an unchanged joint graph.
    // The device colouring validates in its own first kernel, so the host walks the joints only
    // for host-built schedules or when asked to compare with the previous call.
    const bool wantKey = (flags & PHYX_B200_SOLVE_KEEP_SCHEDULE) != 0;
    // Host-built schedules (reference-order replay, cross-check or fallback colouring) and KEEP_SCHEDULE read
    // the joint list on the host: fetch it if the joints were produced on the device.
    auto fetch_joints = [&]() -> int {
        if (hostJoints) return PHYX_B200_OK;
        c->hostJoints.resize(size_t(nj));
        if (nj)
        {
            PHYX_CUDA(cudaMemcpyAsync(c->hostJoints.data(), c->joints.ptr, size_t(nj) * sizeof(phyx_contact_joint), cudaMemcpyDeviceToHost, c->stream));
            PHYX_CUDA(cudaStreamSynchronize(c->stream));
        }
        c->hostJointsValid = true;
        hostJoints = c->hostJoints.data();
        return PHYX_B200_OK;
    };
    if (wantKey || !deviceColouring) PHYX_TRY(fetch_joints());
    std::vector<int> key(wantKey ? size_t(nj) * 2 : 0);
    for (int j = 0; j < nj && (wantKey || !deviceColouring); ++j)
    {
        int b1 = hostJoints[j].body1Index, b2 = hostJoints[j].body2Index, cp = hostJoints[j].contactPointIndex;
        if (b1 < 0 || b1 >= nb || b2 < 0 || b2 >= nb || cp < 0 || cp >= ncp)
        {
            set_error("solve: joint %d references a body outside [0,%d) or a contact point outside [0,%d)", j, nb, ncp);
            return PHYX_B200_ERR_ARGUMENT;
        }
        if (wantKey)
        {
            key[2 * j] = b1;
            key[2 * j + 1] = b2;
        }
    }
    const bool keep = wantKey && c->scheduleMode == mode && c->scheduleFlags == (flags & PHYX_B200_SOLVE_STATIC_DEPS) && key == c->hostPairKey;
    if (keep) return PHYX_B200_OK;
    c->hostPairKey.swap(key);
    c->scheduleMode = mode;
    c->scheduleFlags = flags & PHYX_B200_SOLVE_STATIC_DEPS;

    // throughput schedule: colour on the device (the joints are already resident)
    if (deviceColouring)
    {
        int st = colour_schedule_build(c);
        if (st != PHYX_B200_ERR_CAPACITY) return st;
        // more than 64 colours (a dynamic body with dozens of joints): the host builder has no limit
        PHYX_TRY(fetch_joints());
        c->colourStateValid = false;
    }
    c->hostSlotsStale = false;
    c->hostLevelsStale = false;
    c->strip.valid = false;   // host-built schedules are colour-major / replay levels
    c->colourRounds = 0;

    // static flags from the resident body parameters
    std::vector<float4> params(size_t(nb > 0 ? nb : 1));
    if (nb > 0)
    {
        PHYX_CUDA(cudaMemcpyAsync(params.data(), c->params.ptr, size_t(nb) * sizeof(float4), cudaMemcpyDeviceToHost, c->stream));
        PHYX_CUDA(cudaStreamSynchronize(c->stream));
    }
    std::vector<unsigned char> statics(size_t(nb > 0 ? nb : 1));
    for (int i = 0; i < nb; ++i) statics[i] = is_static(params[i]);

    std::vector<int>& slots = c->hostSlots;
    std::vector<int>& slotPos = c->hostSlotPos;
    std::vector<Level>& levels = c->hostLevels;
    slots.clear();
    slotPos.clear();
    levels.clear();
    // strict companion of a replay schedule (see StaticRule): S position -> slot of the fast schedule
    std::vector<int> strictMap;
    std::vector<Level> strictLevels;
    std::vector<unsigned char> multi;
    int numMulti = 0;
    const int N = mode == PHYX_B200_SCHEDULE_REPLAY_AVX2 ? 8 : mode == PHYX_B200_SCHEDULE_REPLAY_SSE2 ? 4 : 1;
    switch (mode)
    {
    case PHYX_B200_SCHEDULE_COLOUR: build_colours(hostJoints, nj, nb, statics, slots, levels); break;
    case PHYX_B200_SCHEDULE_REPLAY_AVX2:
    case PHYX_B200_SCHEDULE_REPLAY_SSE2:
    case PHYX_B200_SCHEDULE_REPLAY_SCALAR:
    {
        std::vector<int> order;
        std::vector<Unit> units;
        reference_units(hostJoints, nj, nb, N, order, units);
        if (flags & PHYX_B200_SOLVE_STATIC_DEPS)
        {
            layout_replay(hostJoints, nb, statics, order, units, kStaticsSerial, slots, slotPos, levels);
            break;
        }
        // static bodies with at least two units: the only ones whose lastIteration couples joints
        std::vector<int> unitsOn(size_t(nb > 0 ? nb : 1), 0);
        for (const Unit& u : units)
            for (int k = 0; k < u.width; ++k)
            {
                const phyx_contact_joint& j = hostJoints[order[u.first + k]];
                if (statics[j.body1Index]) unitsOn[j.body1Index]++;
                if (statics[j.body2Index]) unitsOn[j.body2Index]++;
            }
        multi.assign(size_t(nb > 0 ? nb : 1), 0);
        for (int b = 0; b < nb; ++b)
            if (unitsOn[b] >= 2)
            {
                multi[b] = 1;
                ++numMulti;
            }
        layout_replay(hostJoints, nb, statics, order, units, kStaticsFree, slots, slotPos, levels);
        if (numMulti > 0)
        {
            std::vector<int> strictSlots, strictPos;
            layout_replay(hostJoints, nb, statics, order, units, kStaticsOrdered, strictSlots, strictPos, strictLevels);
            std::vector<int> slotOfJoint(size_t(nj > 0 ? nj : 1), -1);
            for (size_t k = 0; k < slots.size(); ++k)
                if (slots[k] >= 0) slotOfJoint[slots[k]] = int(k);
            strictMap.resize(strictSlots.size());
            for (size_t k = 0; k < strictSlots.size(); ++k) strictMap[k] = strictSlots[k] >= 0 ? slotOfJoint[strictSlots[k]] : -1;
        }
        break;
    }
    default: set_error("solve: unknown schedule %d", mode); return PHYX_B200_ERR_ARGUMENT;
    }
    c->strictLevelCount = int(strictLevels.size());
    c->numMultiStatics = numMulti;
    if (!strictLevels.empty())
    {
        PHYX_TRY(c->strictLevels.reserve(strictLevels.size() * sizeof(Level)));
        PHYX_TRY(c->strictMap.reserve(strictMap.size() * sizeof(int)));
        PHYX_TRY(c->staticMulti.reserve(multi.size()));
        PHYX_CUDA(cudaMemcpyAsync(c->strictLevels.ptr, strictLevels.data(), strictLevels.size() * sizeof(Level), cudaMemcpyHostToDevice, c->stream));
        PHYX_CUDA(cudaMemcpyAsync(c->strictMap.ptr, strictMap.data(), strictMap.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
        PHYX_CUDA(cudaMemcpyAsync(c->staticMulti.ptr, multi.data(), multi.size(), cudaMemcpyHostToDevice, c->stream));
    }
    c->slotCount = int(slots.size());
    c->levelCount = int(levels.size());
    PHYX_TRY(c->slotJoint.reserve(std::max<size_t>(slots.size(), 1) * sizeof(int)));
    PHYX_TRY(c->levels.reserve(std::max<size_t>(levels.size(), 1) * sizeof(Level)));
    if (!slots.empty())
        PHYX_CUDA(cudaMemcpyAsync(c->slotJoint.ptr, slots.data(), slots.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    if (!levels.empty())
        PHYX_CUDA(cudaMemcpyAsync(c->levels.ptr, levels.data(), levels.size() * sizeof(Level), cudaMemcpyHostToDevice, c->stream));
    // colour schedules run in slot order, so a slot's sequential position is its own index
    c->slotPosValid = !slotPos.empty();
    if (c->slotPosValid)
    {
        PHYX_TRY(c->slotPos.reserve(slotPos.size() * sizeof(int)));
        PHYX_CUDA(cudaMemcpyAsync(c->slotPos.ptr, slotPos.data(), slotPos.size() * sizeof(int), cudaMemcpyHostToDevice, c->stream));
    }
    PHYX_CUDA(cudaStreamSynchronize(c->stream));
    return PHYX_B200_OK;
}

} // namespace phyx
