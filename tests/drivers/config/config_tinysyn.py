# Test harness configuration for the UNMODIFIED reference text.py (`--dataset tinysyn`): the reference's `config` is a
# namespace package, so this directory on sys.path adds `config.config_tinysyn` next to the shipped configs.
# Dropout 0: on the same device the reference back-end and the lagvae back-end then consume torch's generator
# identically (eps only), and the printed trajectories can be compared tightly.
params = {
    'enc_type': 'lstm',
    'dec_type': 'lstm',
    'nz': 8,
    'ni': 64,
    'enc_nh': 256,
    'dec_nh': 256,
    'dec_dropout_in': 0.0,
    'dec_dropout_out': 0.0,
    'batch_size': 16,
    'epochs': 2,
    'test_nepoch': 1,
    'train_data': 'datasets/tinysyn_data/train.txt',
    'val_data': 'datasets/tinysyn_data/valid.txt',
    'test_data': 'datasets/tinysyn_data/test.txt'
}
