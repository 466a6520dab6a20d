//! Drop-in batched environment for `physics_reinforcement_learning_environment`.
//!
//! `World`, `ObjectAndTransform`, `WorldObject` and `Move` are declared exactly as in the
//! reference's src/common.rs:19-49, 351-356 (same serde schema), so inside the reference crate
//! this file's types are simply the crate's own and only `BatchedEnvironment` and
//! `BatchedGeneticAlgorithm` are new. The single-environment `Environment` (rapier on the CPU)
//! stays as it is for the Game and Train viewers (game.rs:126-146, train.rs:132-159).
pub mod ffi;

use serde::{Deserialize, Serialize};
use std::ffi::CStr;
use std::ptr;

#[derive(Serialize, Deserialize, Default, Debug, Clone)]
pub struct World {
    pub player_position: [f32; 2],
    pub objects: Vec<ObjectAndTransform>,
}

#[derive(Serialize, Deserialize, Debug, Clone)]
pub struct ObjectAndTransform {
    pub object: WorldObject,
    pub position: [f32; 3],
    pub scale: [f32; 2],
    pub rotation: f32,
}

#[derive(Serialize, Deserialize, Clone, Debug)]
pub enum WorldObject {
    Block { fixed: bool },
    Goal,
}

#[derive(Default, Clone, Copy, Debug)]
pub struct Move {
    pub left: bool,
    pub right: bool,
    pub up: bool,
}

impl Move {
    /// wire format of the C ABI: bit0 left, bit1 right, bit2 up
    pub fn bits(self) -> u8 {
        (self.left as u8) | ((self.right as u8) << 1) | ((self.up as u8) << 2)
    }
}

#[derive(Debug)]
pub struct PrlError {
    pub status: i32,
    pub message: String,
}

fn check(status: i32) -> Result<(), PrlError> {
    if status == ffi::PRL_OK {
        return Ok(());
    }
    let message = unsafe { CStr::from_ptr(ffi::prl_last_error()) }.to_string_lossy().into_owned();
    Err(PrlError { status, message })
}

struct WorldHandle(*mut ffi::prl_world);
impl WorldHandle {
    fn new(world: &World) -> Result<Self, PrlError> {
        let objs: Vec<ffi::prl_object> = world
            .objects
            .iter()
            .map(|o| ffi::prl_object {
                position: o.position,
                scale: o.scale,
                rotation: o.rotation,
                kind: match o.object {
                    WorldObject::Block { fixed: true } => 0,
                    WorldObject::Block { fixed: false } => 1,
                    WorldObject::Goal => 2,
                },
            })
            .collect();
        let mut h = ptr::null_mut();
        check(unsafe { ffi::prl_world_create(world.player_position.as_ptr(), objs.as_ptr(), objs.len() as i32, &mut h) })?;
        Ok(WorldHandle(h))
    }
}
impl Drop for WorldHandle {
    fn drop(&mut self) {
        unsafe { ffi::prl_world_destroy(self.0) }
    }
}

/// `n` independent `Environment`s of one `World` on one GPU. Mirrors `Environment`'s API
/// (common.rs:162-340) with a leading batch axis.
pub struct BatchedEnvironment {
    _world: WorldHandle,
    batch: *mut ffi::prl_batch,
    n: usize,
}
// one thread at a time per handle; Algorithm::train runs off the main thread (train.rs:80)
unsafe impl Send for BatchedEnvironment {}

impl BatchedEnvironment {
    /// `n` x `Environment::from_world(world)`; the second value is from_world's
    /// `Vec<Option<RigidBodyHandle>>` as body indices.
    pub fn from_world(world: &World, n: usize, device: i32) -> Result<(Self, Vec<Option<usize>>), PrlError> {
        let wh = WorldHandle::new(world)?;
        let mut batch = ptr::null_mut();
        check(unsafe { ffi::prl_batch_create(wh.0, n as i64, device, ptr::null_mut(), &mut batch) })?;
        let handles = (0..world.objects.len() as i32)
            .map(|i| {
                let b = unsafe { ffi::prl_world_object_body(wh.0, i) };
                if b < 0 { None } else { Some(b as usize) }
            })
            .collect();
        Ok((BatchedEnvironment { _world: wh, batch, n }, handles))
    }
    pub fn len(&self) -> usize {
        self.n
    }
    pub fn reset(&mut self) -> Result<(), PrlError> {
        check(unsafe { ffi::prl_batch_reset(self.batch) })
    }
    /// `Environment::step` for every environment.
    pub fn step(&mut self, moves: &[Move]) -> Result<(), PrlError> {
        assert_eq!(moves.len(), self.n);
        let bytes: Vec<u8> = moves.iter().map(|m| m.bits()).collect();
        check(unsafe { ffi::prl_batch_step(self.batch, bytes.as_ptr()) })
    }
    /// `Environment::distance_to_goals` for every environment; Ok(false) = the world has no goal (None).
    pub fn distance_to_goals(&mut self, out: &mut [f32]) -> Result<bool, PrlError> {
        assert_eq!(out.len(), self.n);
        match unsafe { ffi::prl_batch_distance_to_goals(self.batch, out.as_mut_ptr()) } {
            ffi::PRL_ERR_NO_GOAL => Ok(false),
            s => check(s).map(|_| true),
        }
    }
    /// `Environment::won` for every environment.
    pub fn won(&mut self, out: &mut [bool]) -> Result<(), PrlError> {
        assert_eq!(out.len(), self.n);
        let mut bytes = vec![0u8; self.n];
        check(unsafe { ffi::prl_batch_won(self.batch, bytes.as_mut_ptr()) })?;
        for (o, b) in out.iter_mut().zip(bytes) {
            *o = b != 0;
        }
        Ok(())
    }
    /// The GA's `agent_score` closure (main.rs:44-70) for every agent at once.
    pub fn rollout(&mut self, genomes: &[Vec<Move>], repeat_move: usize, number_of_steps: usize) -> Result<Vec<f32>, PrlError> {
        assert_eq!(genomes.len(), self.n);
        let len = genomes.first().map_or(0, |g| g.len());
        let flat: Vec<u8> = genomes.iter().flat_map(|g| g.iter().map(|m| m.bits())).collect();
        let mut score = vec![0f32; self.n];
        check(unsafe {
            ffi::prl_batch_rollout(self.batch, flat.as_ptr(), len as i32, repeat_move as i32, number_of_steps as i32, 1,
                                   score.as_mut_ptr(), ptr::null_mut())
        })?;
        Ok(score)
    }
    /// Replay for the Train-state viewer (train.rs:132-159): the agent's moves are played on a fresh
    /// environment on the GPU in one launch; returns, per step, every body's
    /// (x, y, cos, sin, vx, vy, angvel, sleep timer, sleeping), the distance to the goals and won(),
    /// i.e. what `update_visualization` copies into the sprites frame by frame -- without
    /// re-simulating on the CPU.
    pub fn trajectory(&mut self, moves: &[Move], repeat_move: usize, number_of_steps: usize)
                      -> Result<(Vec<f32>, Vec<f32>, Vec<u8>), PrlError> {
        let flat: Vec<u8> = moves.iter().map(|m| m.bits()).collect();
        let nb = unsafe { ffi::prl_batch_num_bodies(self.batch) } as usize;
        let mut state = vec![0f32; number_of_steps * nb * 9];
        let mut dist = vec![0f32; number_of_steps];
        let mut won = vec![0u8; number_of_steps];
        check(unsafe {
            ffi::prl_batch_trajectory(self.batch, 1, flat.as_ptr(), flat.len() as i32, repeat_move as i32, number_of_steps as i32,
                                      state.as_mut_ptr(), dist.as_mut_ptr(), won.as_mut_ptr())
        })?;
        Ok((state, dist, won))
    }
}
impl Drop for BatchedEnvironment {
    fn drop(&mut self) {
        unsafe { ffi::prl_batch_destroy(self.batch) }
    }
}

/// The reference's `GeneticAlgorithm` (main.rs:18-38) with the population evaluated and bred on
/// the GPU(s). `train` has the body of `Algorithm::train` (algorithm.rs:27) up to the agent type:
/// `examples/batched_main.rs` wraps it into the reference's three traits.
#[derive(PartialEq, Clone)]
pub struct BatchedGeneticAlgorithm {
    pub number_of_steps: usize,
    pub number_of_agents: usize,
    pub repeat_move: usize,
    pub mutation_rate: f32,
    pub keep_best: bool,
    pub seed: u64,
    /// CUDA devices to spread the population over (one slice per entry; peer access between
    /// them is required when there are several -- `prl_ga_create_multi`)
    pub devices: Vec<i32>,
}

impl Default for BatchedGeneticAlgorithm {
    fn default() -> Self {
        BatchedGeneticAlgorithm { number_of_steps: 1000, number_of_agents: 1000, repeat_move: 20, mutation_rate: 0.1, keep_best: false, seed: 0xB200,
                                  devices: vec![0] }
    }
}

/// number of CUDA devices the library can see (0: no GPU -- every compute call fails, there is no CPU path)
pub fn device_count() -> i32 {
    unsafe { ffi::prl_device_count() }
}

impl BatchedGeneticAlgorithm {
    /// all visible devices
    pub fn on_all_devices(mut self) -> Self {
        self.devices = (0..device_count().max(1)).collect();
        self
    }

    /// main.rs:41-163: sends `(best score, best genome)` once per generation until the receiver is
    /// dropped (main.rs:121-124). One `prl_ga_step_generation` call is one iteration of the
    /// reference's `loop` (main.rs:86-162); with several devices the library exchanges the
    /// generation over peer memory itself.
    pub fn train(&self, world: World, sender: crossbeam::channel::Sender<(f32, Vec<Move>)>) -> Result<(), PrlError> {
        let wh = WorldHandle::new(&world)?;
        let cfg = ffi::prl_ga_config {
            number_of_steps: self.number_of_steps as i32,
            number_of_agents: self.number_of_agents as i64,
            repeat_move: self.repeat_move as i32,
            mutation_rate: self.mutation_rate,
            keep_best: self.keep_best as i32,
            seed: self.seed,
        };
        let mut ga = ptr::null_mut();
        check(unsafe { ffi::prl_ga_create_multi(wh.0, &cfg, self.devices.as_ptr(), self.devices.len() as i32, &mut ga) })?;
        struct Guard(*mut ffi::prl_ga);
        impl Drop for Guard {
            fn drop(&mut self) {
                unsafe { ffi::prl_ga_destroy(self.0) }
            }
        }
        let _guard = Guard(ga);
        let genome_len = self.number_of_steps / self.repeat_move;
        loop {
            let mut score = 0f32;
            let mut genome = vec![0u8; genome_len];
            check(unsafe { ffi::prl_ga_step_generation(ga, &mut score, genome.as_mut_ptr(), ptr::null_mut()) })?;
            let moves = genome.iter().map(|b| Move { left: b & 1 != 0, right: b & 2 != 0, up: b & 4 != 0 }).collect();
            if sender.send((score, moves)).is_err() {
                return Ok(()); // receiver dropped: stop training (main.rs:121-124)
            }
        }
    }
}
