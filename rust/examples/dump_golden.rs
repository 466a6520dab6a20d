//! Golden-vector dumper: runs the REAL reference (`physics_reinforcement_learning_environment`
//! = rapier2d 0.17.2 with `enhanced-determinism`, `Cargo.toml:19`) on the cases listed in
//! `tests/golden/request.json` and writes every f32 as its bit pattern to
//! `tests/golden/rapier_dump.json`. `scripts/import_rapier_golden.py` turns that into
//! `tests/golden/rapier_<case>.npz` (`source = "rapier"`), which `tests/test_golden.py` consumes in
//! tolerance mode -- this is what pins the oracle (and with it the CUDA path) to the reference.
//!
//! SOURCE ONLY here (no cargo/rustc in this repository's image). On a machine with Rust:
//!
//!   python scripts/make_golden.py --request          # writes tests/golden/request.json
//!   cd rust && cargo run --release --features reference --example dump_golden -- ../tests/golden
//!   python scripts/import_rapier_golden.py           # -> tests/golden/rapier_*.npz
//!
//! Per case and genome it records (a) the reference's `agent_score` (main.rs:44-70): minimum
//! distance, 1-based step of the first `won()`, steps run with break-on-win; (b) the full
//! trajectory WITHOUT early exit: after every `Environment::step`, for the player and every
//! pushable block in `from_world` handle order: x, y, cos, sin, vx, vy, angvel, sleep timer,
//! sleeping -- the layout of `prl_batch_get_state` -- plus `distance_to_goals()` and `won()`.
use physics_reinforcement_learning_environment::{Environment, Move, World};
use serde::{Deserialize, Serialize};

#[derive(Deserialize)]
struct Request {
    cases: Vec<Case>,
}
#[derive(Deserialize)]
struct Case {
    name: String,
    world: World,
    genomes: Vec<Vec<u8>>,
    repeat_move: usize,
    number_of_steps: usize,
}
#[derive(Serialize)]
struct CaseDump {
    name: String,
    score_bits: Vec<u32>,
    step_of_win: Vec<i32>,
    steps_run: Vec<i32>,
    n_bodies: usize,
    /// [genome][step][body * 9 + field] as f32 bit patterns
    state_bits: Vec<Vec<Vec<u32>>>,
    /// [genome][step]
    distance_bits: Vec<Vec<u32>>,
    won: Vec<Vec<u8>>,
}

fn mv(bits: u8) -> Move {
    Move { left: bits & 1 != 0, right: bits & 2 != 0, up: bits & 4 != 0 }
}

/// main.rs:44-70, verbatim semantics (including the tail that runs one step after a win)
fn agent_score(world: &World, genome: &[u8], repeat_move: usize, number_of_steps: usize) -> (f32, i32, i32) {
    let (mut env, _) = Environment::from_world(world);
    let (mut score, mut win, mut steps) = (f32::INFINITY, -1i32, 0i32);
    let mut one = |env: &mut Environment, m: Move| {
        env.step(m);
        steps += 1;
        score = score.min(env.distance_to_goals().unwrap());
        if env.won() && win < 0 {
            win = steps;
        }
        env.won()
    };
    'genes: for &g in genome {
        for _ in 0..repeat_move {
            if one(&mut env, mv(g)) {
                break 'genes;
            }
        }
    }
    for _ in 0..number_of_steps % repeat_move {
        if one(&mut env, Move::default()) {
            break;
        }
    }
    (score, win, steps)
}

fn main() {
    let dir = std::env::args().nth(1).unwrap_or_else(|| "../tests/golden".into());
    let request: Request = serde_json::from_str(&std::fs::read_to_string(format!("{dir}/request.json")).unwrap()).unwrap();
    let mut dumps = Vec::new();
    for case in &request.cases {
        let mut d = CaseDump { name: case.name.clone(), score_bits: vec![], step_of_win: vec![], steps_run: vec![], n_bodies: 0,
                               state_bits: vec![], distance_bits: vec![], won: vec![] };
        for genome in &case.genomes {
            let (score, win, steps) = agent_score(&case.world, genome, case.repeat_move, case.number_of_steps);
            d.score_bits.push(score.to_bits());
            d.step_of_win.push(win);
            d.steps_run.push(steps);
            // full trajectory, GeneticAgent::get_move semantics (main.rs:276-284), no early exit
            let (mut env, handles) = Environment::from_world(&case.world);
            let mut bodies = vec![env.player_handle()];
            bodies.extend(handles.iter().flatten().copied());
            d.n_bodies = bodies.len();
            let (mut states, mut dists, mut wons) = (vec![], vec![], vec![]);
            for t in 0..case.number_of_steps {
                let gene = genome.get(t / case.repeat_move).copied().unwrap_or(0);
                env.step(mv(gene));
                let mut row = Vec::with_capacity(bodies.len() * 9);
                for h in &bodies {
                    let rb = &env.rigid_body_set()[*h];
                    let f = [rb.translation().x, rb.translation().y, rb.rotation().re, rb.rotation().im, rb.linvel().x, rb.linvel().y,
                             rb.angvel(), rb.activation().time_since_can_sleep, if rb.is_sleeping() { 1.0 } else { 0.0 }];
                    row.extend(f.iter().map(|v| v.to_bits()));
                }
                states.push(row);
                dists.push(env.distance_to_goals().unwrap_or(f32::INFINITY).to_bits());
                wons.push(env.won() as u8);
            }
            d.state_bits.push(states);
            d.distance_bits.push(dists);
            d.won.push(wons);
        }
        dumps.push(d);
    }
    std::fs::write(format!("{dir}/rapier_dump.json"), serde_json::to_string(&dumps).unwrap()).unwrap();
    eprintln!("wrote {} cases to {dir}/rapier_dump.json", dumps.len());
}
