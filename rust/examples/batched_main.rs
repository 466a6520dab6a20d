//! Drop-in replacement for the reference's `src/main.rs`: the same three plug-in types
//! (`GeneticAgent`, `GeneticTrainingDetails`, `GeneticAlgorithm`) implementing the reference's own
//! traits (`src/algorithm.rs:8-33`), handed to the reference's `run::<..>()` (`src/lib.rs:172-189`),
//! with `Algorithm::train` running on the GPU batch instead of one rapier `Environment` per agent.
//!
//! SOURCE ONLY: this image has no cargo/rustc. Build on a machine with a Rust toolchain and the
//! reference checked out beside this repository:
//!
//!   cd rust && RUSTFLAGS="-L ../physics_reinforcement_learning_environment_b200 -l prl" \
//!       cargo run --release --features reference --example batched_main
//!
//! The viewers are untouched: a trained agent is still replayed by `update_visualization`
//! (train.rs:132-159) on the CPU `Environment`, through `Agent::get_move` below.
use physics_reinforcement_learning_environment::{
    egui::{self, Ui},
    Agent, Algorithm, Environment, Move, Receiver, Sender, TrainingDetails, World,
};
use prl_b200::{device_count, BatchedGeneticAlgorithm};

fn main() {
    physics_reinforcement_learning_environment::run::<GeneticAgent, GeneticMessage, GeneticTrainingDetails, GeneticAlgorithm>();
}

type GeneticMessage = (f32, GeneticAgent);

/// Fields and defaults of the reference's `GeneticAlgorithm` (main.rs:18-38) plus the device count.
/// The population is no longer capped at the UI's 1000 agents (main.rs:170-173): the batch scores
/// a million agents per generation in a fraction of a second.
#[derive(PartialEq, Clone)]
pub struct GeneticAlgorithm {
    number_of_steps: usize,
    number_of_agents: usize,
    repeat_move: usize,
    mutation_rate: f32,
    keep_best: bool,
    gpus: usize,
}

impl Default for GeneticAlgorithm {
    fn default() -> Self {
        GeneticAlgorithm { number_of_steps: 1000, number_of_agents: 1000, repeat_move: 20, mutation_rate: 0.1, keep_best: false, gpus: 1 }
    }
}

impl Algorithm<GeneticAgent, GeneticMessage, GeneticTrainingDetails> for GeneticAlgorithm {
    fn selection_ui(&mut self, ui: &mut Ui) {
        ui.add(egui::Slider::new(&mut self.number_of_steps, 1..=100_000).text("steps per rollout"));
        ui.add(egui::Slider::new(&mut self.number_of_agents, 10..=4_000_000).logarithmic(true).text("agents"));
        ui.add(egui::Slider::new(&mut self.repeat_move, 1..=100).text("steps per gene"));
        ui.add(egui::Slider::new(&mut self.mutation_rate, 0.0..=1.0).text("mutation rate"));
        ui.checkbox(&mut self.keep_best, "carry the best agent over");
        let visible = device_count().max(1) as usize;
        ui.add(egui::Slider::new(&mut self.gpus, 1..=visible).text("GPUs"));
    }

    /// algorithm.rs:27. The whole generation loop (main.rs:72-162) runs behind the C ABI; this
    /// thread only forwards one message per generation and stops when the UI drops the receiver
    /// (train.rs:89-92, main.rs:121-124).
    fn train(&self, world: World, sender: Sender<GeneticMessage>) {
        let algo = BatchedGeneticAlgorithm {
            number_of_steps: self.number_of_steps,
            number_of_agents: self.number_of_agents,
            repeat_move: self.repeat_move,
            mutation_rate: self.mutation_rate,
            keep_best: self.keep_best,
            devices: (0..self.gpus as i32).collect(),
            ..Default::default()
        };
        // the reference's World and the facade's World share one serde schema (common.rs:19-49)
        let world: prl_b200::World = serde_json::from_str(&serde_json::to_string(&world).unwrap()).unwrap();
        let repeat_move = self.repeat_move;
        let (tx, rx) = crossbeam::channel::bounded::<(f32, Vec<prl_b200::Move>)>(1);
        std::thread::scope(|scope| {
            scope.spawn(move || {
                if let Err(e) = algo.train(world, tx) {
                    eprintln!("batched GA stopped: {} ({})", e.message, e.status);
                }
            });
            for (score, genome) in rx {
                let moves = genome.iter().map(|m| Move { left: m.left, right: m.right, up: m.up }).collect();
                if sender.send((score, GeneticAgent { moves, curr: 0, repeat_move })).is_err() {
                    break; // dropping rx ends the device loop at its next send
                }
            }
        });
    }

    fn training_details_receiver(&self, _world: &World, receiver: Receiver<GeneticMessage>) -> GeneticTrainingDetails {
        GeneticTrainingDetails { generations: Vec::new(), receiver }
    }
}

pub struct GeneticTrainingDetails {
    generations: Vec<GeneticMessage>,
    receiver: Receiver<GeneticMessage>,
}

impl TrainingDetails<GeneticAgent, GeneticMessage> for GeneticTrainingDetails {
    fn receive_messages(&mut self) {
        while let Ok(msg) = self.receiver.try_recv() {
            self.generations.push(msg);
        }
    }

    fn details_ui(&mut self, ui: &mut Ui) -> Option<&GeneticAgent> {
        let mut chosen = None;
        for (generation, (score, agent)) in self.generations.iter().enumerate() {
            if ui.button(format!("generation {generation}: distance {score:.2} -- replay")).clicked() {
                chosen = Some(agent);
            }
        }
        chosen
    }
}

/// main.rs:223-284: a fixed move list, each move held for `repeat_move` steps.
#[derive(Clone)]
pub struct GeneticAgent {
    moves: Vec<Move>,
    curr: usize,
    repeat_move: usize,
}

impl Agent for GeneticAgent {
    fn get_move(&mut self, _environment: &Environment) -> Move {
        match self.moves.get(self.curr / self.repeat_move) {
            Some(&mv) => {
                self.curr += 1;
                mv
            }
            None => Move::default(),
        }
    }

    fn details_ui(&self, ui: &mut Ui, _environment: &Environment) {
        let at = self.curr / self.repeat_move;
        ui.label(format!("gene {} of {}, {} steps each", (at + 1).min(self.moves.len()), self.moves.len(), self.repeat_move));
        let text: String = self.moves.iter().map(|m| match (m.left, m.right, m.up) {
            (false, false, false) => '.', (true, false, false) => '<', (false, true, false) => '>', (false, false, true) => '^',
            (true, false, true) => '\\', (false, true, true) => '/', (true, true, false) => '=', (true, true, true) => '#',
        }).collect();
        ui.monospace(text);
    }
}
