#!/usr/bin/env python3
"""``compare`` / ``search`` / ``encrypt-batch`` with the reference CLI's argument surface and output
(/root/reference/fhe_cli.py:106-210,327-346), plus ``keys generate`` (fhe_cli.py:42-60) for the one piece of key
management the encrypted-compare path needs: a persistent key set.  The reference's key manager stores
a config dict behind a master password (key_management.py:148-166); here the file holds the seeds every key is derived
from, under the same PBKDF2 + Fernet wrapper.  The master password comes from ``FHE_MASTER_PASSWORD`` or, like the
reference (key_management.py:69,87), from ``getpass``.  stats / validate / estimate / key rotation are outside the path
and not provided."""
from __future__ import annotations

import argparse
import json
import os
import sys
from pathlib import Path

from .batch_operations import BatchProcessor, DocumentStore


def _master_password(confirm: bool = False) -> str:
    pw = os.environ.get("FHE_MASTER_PASSWORD")
    if pw is not None:
        return pw
    import getpass
    pw = getpass.getpass("Enter master password: ")
    if confirm and getpass.getpass("Confirm master password: ") != pw:
        raise ValueError("Passwords don't match")
    return pw


class FHEDocumentCLI:
    def __init__(self, storage_dir: str = "./encrypted_docs", fhe: str = "execute", seed: int = 0, keys_path: str = None):
        self.storage = DocumentStore(storage_dir)
        self.fhe, self.seed = fhe, seed
        # fhe="both" keeps ciphertexts at rest, so the key set must outlive the process: default key file in the store
        self.keys_path = keys_path or (str(Path(storage_dir) / "keys.fhe") if fhe == "both" else None)
        self._processor = None

    def _load_or_create_keys(self):
        from .serialization import KeySet, load_keys, save_keys
        if self.keys_path is None:
            return None                                   # ephemeral key set (fresh from the OS CSPRNG)
        if Path(self.keys_path).exists():
            return load_keys(self.keys_path, _master_password())
        keys = KeySet.generate()
        Path(self.keys_path).parent.mkdir(parents=True, exist_ok=True)
        save_keys(self.keys_path, keys, _master_password(confirm=True))
        print(f"Generated a new key set: {self.keys_path}")
        return keys

    def _get_processor(self) -> BatchProcessor:
        if self._processor is None:
            self._processor = BatchProcessor(storage=self.storage, fhe=self.fhe, seed=self.seed,
                                             keys=self._load_or_create_keys())
        return self._processor

    def cmd_keys(self, args):
        from .serialization import KeySet, save_keys
        if args.key_command != 'generate':
            print("Only `keys generate` is provided (rotation / listing are outside the encrypted-compare path)")
            return
        if self.keys_path is None:
            self.keys_path = str(Path(self.storage.dir or ".") / "keys.fhe")
        if Path(self.keys_path).exists() and not args.force:
            print(f"Error: {self.keys_path} exists (use --force to overwrite: stored ciphertexts become undecryptable)")
            return
        Path(self.keys_path).parent.mkdir(parents=True, exist_ok=True)
        save_keys(self.keys_path, KeySet.generate(), _master_password(confirm=True))
        print("Generating FHE keys...")
        print(f"Keys generated successfully!\nKey file: {self.keys_path}")

    def cmd_encrypt_batch(self, args):
        processor = self._get_processor()
        with open(args.input_file, 'r') as f:
            data = json.load(f)
        if not isinstance(data, list):
            print("Error: Input file must contain a JSON array of documents")
            return
        texts, doc_ids, metadata_list = [], [], []
        for item in data:
            if isinstance(item, str):
                texts.append(item); doc_ids.append(None); metadata_list.append({})
            elif isinstance(item, dict):
                texts.append(item.get('text', '')); doc_ids.append(item.get('id')); metadata_list.append(item.get('metadata', {}))
            else:
                print(f"Warning: Skipping invalid item: {item}")
        print(f"Encrypting {len(texts)} documents...")
        encrypted_ids = processor.encrypt_documents(texts, doc_ids, metadata_list)
        print(f"\nEncrypted {len(encrypted_ids)} documents successfully!")
        if args.output_file:
            with open(args.output_file, 'w') as f:
                json.dump(encrypted_ids, f, indent=2)
            print(f"Document IDs saved to: {args.output_file}")

    def cmd_compare(self, args):
        processor = self._get_processor()
        print("Comparing documents...")
        print(f"  Document 1: {args.doc1}")
        print(f"  Document 2: {args.doc2}")
        try:
            similarity = processor.compare_encrypted(args.doc1, args.doc2)
            print(f"\nSimilarity score: {similarity:.4f}")
            if similarity > 0.9:
                interpretation = "Very similar"
            elif similarity > 0.7:
                interpretation = "Similar"
            elif similarity > 0.5:
                interpretation = "Somewhat similar"
            else:
                interpretation = "Not very similar"
            print(f"Interpretation: {interpretation}")
        except Exception as e:
            print(f"Error: {e}")

    def cmd_search(self, args):
        processor = self._get_processor()
        print(f"Searching for documents similar to: '{args.query}'")
        print(f"Top {args.top_k} results with similarity >= {args.min_similarity}")
        results = processor.search_similar(args.query, top_k=args.top_k, min_similarity=args.min_similarity)
        if not results:
            print("\nNo similar documents found.")
            return
        print(f"\nFound {len(results)} similar documents:")
        for i, (doc_id, score) in enumerate(results, 1):
            doc_info = self.storage.index.get(doc_id, {})
            print(f"\n{i}. {doc_id} (similarity: {score:.4f})")
            if doc_info.get('metadata'):
                print(f"   Metadata: {doc_info['metadata']}")


def build_parser() -> argparse.ArgumentParser:
    parser = argparse.ArgumentParser(description="FHE Document Encryption and Comparison CLI (B200 engine)")
    parser.add_argument('--storage-dir', default='./encrypted_docs')
    parser.add_argument('--fhe', default='execute', choices=['execute', 'disable', 'both'],
                        help="execute = encrypted evaluation on the GPU (default); disable = the reference's clear path; "
                             "both = both vectors encrypted, products evaluated by programmable bootstraps")
    parser.add_argument('--seed', type=int, default=0)
    parser.add_argument('--keys', default=None, metavar='FILE',
                        help="key file (created on first use; master password from FHE_MASTER_PASSWORD or a prompt). "
                             "Default: <storage-dir>/keys.fhe for --fhe both, an ephemeral key set otherwise")
    subparsers = parser.add_subparsers(dest='command', help='Available commands')
    keys_parser = subparsers.add_parser('keys', help='Key management')
    keys_sub = keys_parser.add_subparsers(dest='key_command')
    gen = keys_sub.add_parser('generate', help='Generate new keys')
    gen.add_argument('--force', action='store_true')
    batch_parser = subparsers.add_parser('encrypt-batch', help='Encrypt multiple documents')
    batch_parser.add_argument('input_file', help='JSON file with documents')
    batch_parser.add_argument('--output-file', '-o', help='Save IDs to file')
    compare_parser = subparsers.add_parser('compare', help='Compare two documents')
    compare_parser.add_argument('doc1', help='First document ID')
    compare_parser.add_argument('doc2', help='Second document ID')
    search_parser = subparsers.add_parser('search', help='Search for similar documents')
    search_parser.add_argument('query', help='Query text')
    search_parser.add_argument('--top-k', type=int, default=5, help='Number of results (default: 5)')
    search_parser.add_argument('--min-similarity', type=float, default=0.5, help='Minimum similarity (default: 0.5)')
    return parser


def main(argv=None):
    parser = build_parser()
    args = parser.parse_args(argv)
    if not args.command:
        parser.print_help()
        return 0
    cli = FHEDocumentCLI(args.storage_dir, args.fhe, args.seed, args.keys)
    handler = {'encrypt-batch': cli.cmd_encrypt_batch, 'compare': cli.cmd_compare, 'search': cli.cmd_search,
               'keys': cli.cmd_keys}[args.command]
    try:
        handler(args)
    except KeyboardInterrupt:
        print("\nOperation cancelled.")
    except Exception as e:  # same behaviour as the reference: log and exit 1
        print(f"Error: {e}", file=sys.stderr)
        return 1
    return 0


if __name__ == "__main__":
    sys.exit(main())
