# Test harness configuration for the UNMODIFIED reference text.py (`--dataset tinysyndrop`): the reference's `config` is a
# namespace package, so this directory on sys.path adds `config.config_tinysyndrop` next to the shipped configs.
# Dropout 0.5 (the reference default): masks differ between the back-ends (torch bernoulli vs in-kernel Philox),
# so the trajectories are compared statistically (loose tolerance).
params = {
    'enc_type': 'lstm',
    'dec_type': 'lstm',
    'nz': 8,
    'ni': 64,
    'enc_nh': 256,
    'dec_nh': 256,
    'dec_dropout_in': 0.5,
    'dec_dropout_out': 0.5,
    'batch_size': 16,
    'epochs': 2,
    'test_nepoch': 1,
    'train_data': 'datasets/tinysyn_data/train.txt',
    'val_data': 'datasets/tinysyn_data/valid.txt',
    'test_data': 'datasets/tinysyn_data/test.txt'
}
